#!/usr/bin/env bash
# round 2, GPU call 26: last check of the committed tree: smoke, whole GPU suite, a short bench line (sharded block at N=1 included), topology
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_c26_topo.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c26_smoke.txt 2>&1; tail -1 gpurun_out/r02_c26_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_c26_tests.txt; tail -2 gpurun_out/r02_c26_tests.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r02_c26_bench.json 2> gpurun_out/r02_c26_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_c26_bench.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), 'sharded:', json.dumps(d['sharded'])[:300])
PY
