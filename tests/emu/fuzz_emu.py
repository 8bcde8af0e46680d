#!/usr/bin/env python
"""Differential fuzz: host build of the engine's device functions (tests/emu) vs the pinned oracle
(oracle/horus_oracle.c) on random clusters / traces / schedules (horus, gandiva, horus+), both stream forms
(standard-normal values, raw MT19937 words) and resumed runs.   python tests/emu/fuzz_emu.py FIRST LAST
Not part of the suite (tests/test_horus_emu.py holds a fixed subset); last run: DESIGN.md section 5."""
import sys, importlib.util, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
spec = importlib.util.spec_from_file_location("tests_emu", "/root/repo/tests/emu/__init__.py")
mod = importlib.util.module_from_spec(spec); sys.modules["tests_emu"]=mod; spec.loader.exec_module(mod)
import numpy as np, oracle
from gpuschedule_b200 import capi, ingest, tracegen
from tests_emu import run_horus
bad=0; t0=time.time()
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng=np.random.default_rng(50000+seed)
    kind=str(rng.choice(["horus","gandiva","horus+"]))
    scheme=str(rng.choice(["horus","gandiva","horus+","yarn"]))      # placement routine; the score follows the schedule
    G=int(rng.choice([2,4,8])); gpc=int(rng.choice([1,1,2]))
    cluster=capi.make_cluster(num_switch=int(rng.integers(1,4)),num_node_p_switch=int(rng.integers(1,6)),num_gpu_p_node=G,
        num_cpu_p_node=int(rng.choice([36,60,128])),mem_p_node=int(rng.choice([180,300,512])),gpu_memory_capacity=int(rng.choice([16,32])))
    choices=sorted(set(int(x)*gpc for x in rng.choice([1,1,2,3,4,6,8,12],size=4)))
    table=ingest.table_from_columns(tracegen.synth_columns(int(rng.integers(10,160)),seed=60000+seed,rate=float(rng.choice([0.5,1,2,4])),
        gpu_per_container=gpc,gpu_choices=choices,gpu_probs=rng.dirichlet(np.ones(len(choices))),max_mem_mib=int(rng.choice([6000,16384,33500]))))
    params=dict(scheme=scheme,schedule=kind,num_buffer=int(rng.choice([1,2,5,15])),num_queue=int(rng.integers(1,7)) if kind=="horus+" else 1,seed=int(rng.integers(0,2**31-1)))
    ref=oracle.run_horus(cluster,table,**params)
    hp=capi.make_horus_params(scheme,kind,params["num_buffer"],params["num_queue"])
    np.random.seed(params["seed"])
    need_words=int(ref.draws*3+200000+ref.ticks*50)
    if kind=="horus+" or seed%2:
        out=run_horus(cluster,hp,table,None,1<<16,int(rng.choice([0,23])),words=np.random.randint(0,2**32,size=need_words,dtype=np.uint32),cooperative=bool(seed%3!=1))
    else:
        out=run_horus(cluster,hp,table,np.random.standard_normal(ref.draws+10),1<<16,int(rng.choice([0,23])),cooperative=bool(seed%4<2))
    ok = out[0]==ref.ticks and out[1].tobytes()==ref.rows.tobytes() and out[2].tobytes()==ref.util.tobytes() and out[3].tobytes()==ref.util_is_array.tobytes() and out[4].tobytes()==ref.recs.tobytes() and np.array_equal(out[5],ref.finish_order) and out[7]==ref.draws
    if not ok: bad+=1; print("MISMATCH",seed,kind,params,out[0],ref.ticks,flush=True)
print("done",int(sys.argv[2])-int(sys.argv[1]),"cases, mismatches",bad,"in %.0fs"%(time.time()-t0))
