"""Extended differential run (not collected by pytest): oracle/tight2_cpu.c -- the algorithm of the CUDA fifo engine as
scalar C, records decoded with the package decoders -- against the pinned literal oracle on random clusters / traces,
whole runs and small resumable windows.   python tests/fuzz_tight2.py FIRST_SEED LAST_SEED
Round 2: seeds 10000-11999 and 20000-39999 (22 000 cases), 0 mismatches."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, oracle
from test_cpu_differential import _case
from test_tight2_cpu import _same
bad=0; t0=time.time(); n=0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    try:
        cluster, table = _case(seed)
        ref = oracle.run_fifo(cluster, table)
        t2 = oracle.Tight2(cluster, table)
        _same(ref, t2.run_all(), f"seed {seed}")
        _same(ref, t2.run_all(max_ticks=1 + seed % 13, cap_a=1 + seed % 5, cap_b=1 + seed % 3), f"seed {seed} windows")
        n+=1
    except AssertionError as e:
        bad+=1; print('MISMATCH', seed, str(e)[:200], flush=True)
    except Exception as e:
        print('skip', seed, repr(e)[:120], flush=True)
print('done', n, 'cases', bad, 'mismatches', '%.0f s'%(time.time()-t0))
