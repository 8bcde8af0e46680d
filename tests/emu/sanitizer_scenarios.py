"""Scenarios for tests/emu/run_sanitizers.sh: every fixture under four mappings / stream forms and the GPU tests' 40
seeded replicas, against the host build of gs_horus.cu given as argv[1] (built with -fsanitize=...)."""
import sys, ctypes as C, functools
import os
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import numpy as np
from gpuschedule_b200 import capi
lib = capi.declare_horus_prototypes(C.CDLL(sys.argv[1]))
import test_horus_abi_emu as T
class cls(capi.HorusEngine):           # test infrastructure: the sanitizer build of the host-emulation library
    @staticmethod
    def _library():
        return lib
from conftest import horus_cases, load_horus, render_horus_outputs
for lanes, words, mt in ((1,False,0),(0,False,0),(0,True,64),(32,False,50)):
    cases = horus_cases(); loaded=[load_horus(c) for c in cases]
    res,_ = T._run(cls, [(cl,tb,pr) for tb,cl,pr,_,_ in loaded], max_ticks=mt, lanes=lanes, words=words)
    ok = all(render_horus_outputs(tb,cl,r)[1]==cc for (tb,cl,pr,jc,cc),r in zip(loaded,res))
    print("lanes",lanes,"words",words,"ok",ok, flush=True)
import test_gpu_widen_horus as G
jobs=[G._seeded_case(s) for s in range(40)]
res,_=T._run(cls, jobs, lanes=0); print("seeded coop done", flush=True)
res,_=T._run(cls, jobs, lanes=1); print("seeded scalar done", flush=True)
