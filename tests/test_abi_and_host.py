"""CPU-side checks: the C-ABI library exports what include/gsched.h declares, fails
loudly without a GPU, and the host logic (flags, ingest, formatting) behaves."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import REPO


def _declared_symbols():
    names = set()
    for header in ("gsched.h", "gsched_horus.h"):
        src = open(os.path.join(REPO, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gs_[a-z_0-9]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from gpuschedule_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert "gs_run" in names and "gs_place_batch" in names and "gs_horus_run" in names and len(names) >= 22
    for name in names:
        assert hasattr(lib, name), name
    assert lib.gs_abi_version() == 4


def test_library_is_built_for_sm_100a():
    from gpuschedule_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    from gpuschedule_b200 import capi
    with pytest.raises(capi.GsError, match="no usable CUDA device|CUDA"):
        capi.Engine(device=0, nsims=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "gpuschedule_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "liboracle" not in text, f
    for f in ("run_sim.py", "execute.py"):
        p = os.path.join(REPO, f)
        if os.path.exists(p):
            assert "oracle" not in open(p).read(), f


def test_struct_sizes_match_header():
    from gpuschedule_b200 import capi, log_manager
    assert ctypes.sizeof(capi.GsCluster) == 56
    assert ctypes.sizeof(capi.GsRunStats) == 72
    assert log_manager.ROW_DTYPE.itemsize == 64
    assert log_manager.JOB_DTYPE.itemsize == 24
    assert capi.NODE_DTYPE.itemsize == 16 and capi.JOBREQ_DTYPE.itemsize == 16


def test_tracegen_is_deterministic_and_in_schema():
    from gpuschedule_b200 import ingest, tracegen
    a = tracegen.synth_columns(500, seed=3)
    b = tracegen.synth_columns(500, seed=3)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    for k in ingest.REQUIRED:
        assert k in a
    t = ingest.table_from_columns(a)
    assert t.n == 500 and np.all(np.diff(t.arrive_tick) >= 0)
    assert np.all(t.gpus % t.gpu_per_task == 0)


def test_ingest_filters_sorts_and_validates(tmp_path):
    import pandas as pd
    from gpuschedule_b200 import ingest, tracegen
    df = tracegen.synth_frame(30, seed=9)
    df.loc[3, "type"] = "interactive"
    df.loc[5, "minutes"] = np.nan
    p = tmp_path / "t.csv"
    df.sample(frac=1.0, random_state=1).to_csv(p, index=False)
    t = ingest.JobTraceReader(str(p)).prepare_jobs().table()
    assert t.n == 28 and len(set(t.label)) == 28
    assert t.arrive_tick[0] == 0 and np.all(np.diff(t.arrive_tick) >= 0)
    bad = tracegen.synth_frame(5, seed=1)
    bad.loc[2, "used_gpus"] = 0
    with pytest.raises(ValueError):
        ingest.table_from_frame(bad.assign(normalized_time=bad["normalized_time"] / 10000))
    with pytest.raises(SystemExit):
        ingest.JobTraceReader(str(tmp_path / "missing.csv"))


def test_flags_surface():
    from gpuschedule_b200 import flags
    F = flags.define_simulator_flags()
    F.reset(["--num_switch", "4", "--enable_network_costs", "False", "--schedule", "fifo", "--nopack"])
    assert F.num_switch == 4 and F.enable_network_costs is False and F.pack is False
    assert F.num_gpu_p_node == 8 and F.mem_p_node == 512 and F.bandwidth == 1250
    F.reset(["--enable_network_costs"])
    assert F.enable_network_costs is True


def test_log_manager_headers_and_object_api(tmp_path):
    from gpuschedule_b200 import log_manager

    class Infra:
        nodes = {str(i): None for i in range(3)}

        def get_total_gpus(self):
            return 6

    class F:
        scheme = "yarn"

    lm = log_manager.LogManager(str(tmp_path), F())
    lm.init(Infra())
    assert open(tmp_path / "cpu.csv", newline="").read() == "time,cpu0,cpu1,cpu2\r\n"
    assert open(tmp_path / "gpu.csv", newline="").read() == "time,gpu0,gpu1,gpu2,gpu3,gpu4,gpu5\r\n"
    assert open(tmp_path / "network.csv", newline="").read() == "time,in0,out0,in1,out1,in2,out2\r\n"
    assert open(tmp_path / "memory.csv", newline="").read() == "time,max,99th,95th,med\r\n"
    lm.step_cluster(log_manager.LogInfo(3, 0, 0, 6, 0.0, 0.0, 0.0, float("nan"), 0, 0, 0, 0), 1)
    lines = open(tmp_path / "cluster.csv", newline="").read().split("\r\n")
    assert lines[1] == "1,3,0,0,6,0.0,0.0,0.0,nan,0,0,0,0"
