#!/bin/bash
# Few-second check of the host-side additions on a real GPU, without pytest/torch start-up:
# examples/octant_cycle.cpp built against the drop-in headers + the CUDA library must print the
# committed golden lines, and the name queries / writers must equal the oracle's.
#   gpurun --timeout 40 -- 'bash tools/gpu/quick_dropin.sh > gpurun_out/quick_dropin.log 2>&1'
set -u
cd "$(dirname "$0")/../.."
H=tmr_b200/csrc/host
g++ -std=c++14 -O1 -I$H -I$H/shim -Iinclude examples/octant_cycle.cpp -Ltmr_b200/lib -ltmr_b200 \
    -Wl,-rpath,$PWD/tmr_b200/lib -pthread -o /tmp/octant_cycle || exit 1
for o in 2 3; do
  /tmp/octant_cycle $o | diff - tests/golden/octant_cycle_order$o.txt > /dev/null \
    && echo "octant_cycle order $o: same lines as the reference" || echo "octant_cycle order $o: DIFFERS"
done
python - <<'PY'
import sys, random
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import util, tmr_b200
from tmr_b200.forest import OctForest
from oracle import ref_loader
gpu, ref = tmr_b200.require_gpu(), ref_loader.load()
conn = util.box_conn()
xpts = np.random.default_rng(1).normal(size=(int(conn.max()) + 1, 3))
res = []
for tag, lib in (("ref", ref), ("gpu", gpu)):
    f = OctForest(order=3, lib=lib); f.setTrilinearTopology(conn, xpts)
    c = f.getConnectivity(); r = random.Random(7)
    for kind, cnt in ((0, c["nnodes"]), (1, c["nedges"]), (2, c["nfaces"]), (3, c["nblocks"])):
        for i in range(cnt):
            n = r.choice([None, "a", "b"])
            if n: f.setEntityName(kind, i, n)
    f.createTrees(1)
    for p in range(2):
        f.refine(util.synth_flags(f.getOctants().as_array(), 5 + p, 30)); f.balance(1)
    f.createNodes()
    f.writeForestToVTK("/tmp/forest_%s.vtk" % tag)
    res.append([f.getOctsWithName(n) for n in ("a", "b")] + [f.getNodesWithName(n) for n in ("a", "b")])
ok = all(np.array_equal(x, y) for x, y in zip(*res))
same_file = open("/tmp/forest_ref.vtk", "rb").read() == open("/tmp/forest_gpu.vtk", "rb").read()
print("name queries:", "equal" if ok else "DIFFER", [len(x) for x in res[1]])
print("writeForestToVTK:", "same bytes" if same_file else "DIFFERS")
PY
