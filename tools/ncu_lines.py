#!/usr/bin/env python
"""Join an ncu report's SASS page with nvdisasm line info and aggregate per CUDA source line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep gs_lane_kernelIj [--top 40]

Needs the .so the report was captured with (same build) at gpuschedule_b200/libgsched.so.
"""
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("NCU_SO", os.path.join(repo, "gpuschedule_b200", "libgsched.so"))
src_dir = os.environ.get("NCU_SRC", os.path.join(repo, "gpuschedule_b200", "csrc"))
_files = {}


def src_line(fname, line):
    """text of `line` in source file `fname` (looked up by base name under the csrc directory)."""
    base = os.path.basename(fname)
    if base not in _files:
        path = os.path.join(src_dir, base) if os.path.isdir(src_dir) else src_dir
        try:
            _files[base] = open(path).read().split("\n")
        except OSError:
            _files[base] = []
    src = _files[base]
    return src[line - 1].strip()[:95] if 0 < line <= len(src) else "?"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = "\n".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
                for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))      # every translation unit's cubin
off2line, cur, infn = {}, ("", 0), False
for ln in dis.split("\n"):
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ie, te, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
base = None
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= te:
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    line = off2line.get(addr - base, ("", -1))
    a = agg.setdefault(line, [0, 0, 0])
    a[0] += int(r[ie]); a[1] += int(r[te]); a[2] += int(r[sm])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot}, samples {tots}")
key = (lambda kv: -kv[1][0]) if "--by-inst" in sys.argv else (lambda kv: -kv[1][2])
for (fname, line), a in sorted(agg.items(), key=key)[:top]:
    print("%s:%d: inst %5.1f%%  samples %5.1f%%  thr/inst %5.1f | %s" % (os.path.basename(fname), line, 100 * a[0] / tot,
                                                                  100 * a[2] / max(tots, 1), a[1] / max(a[0], 1), src_line(fname, line)))
