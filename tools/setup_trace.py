import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from parallelfdtd_b200 import capi, synth
import bench
bid_np, mat_np = synth.shoebox((512,512,512), 6)
bid,_k1 = bench.pinned_u8(bid_np.shape); mat,_k2 = bench.pinned_u8(mat_np.shape); bid[...] = bid_np; mat[...] = mat_np
lam = float(np.sqrt(1/3)); prm = np.array([lam, lam*lam, 1/3, 0], dtype=np.float32)
for dif in (2, 0, 2, 0):
    tab = (synth.filter_material_table(list(np.linspace(0.99,0.5,6)), dif) if dif else synth.material_table(list(np.linspace(0.99,0.5,6)))).astype(np.float32)
    t0 = time.time()
    s = capi.Solver(); s.set_option(capi.OPT_MATIDX_AS_WRITTEN, 0); s.set_option(capi.OPT_DIF_ORDER, dif)
    s.setup_mesh(bid, mat, (32,4,1), 0, capi.F32, prm, tab); t1 = time.time()
    s.make_partition(1, [0]); t2 = time.time()
    src = np.zeros((1, 20), dtype=np.float32); src[0, 1] = 1
    s.set_sources([[256,256,256]], [0], src); s.set_receivers([[270,260,259]])
    r, _ = s.run(20); t3 = time.time()
    s.close(); t4 = time.time()
    print(f"dif {dif}: setup_mesh {1e3*(t1-t0):.1f} ms, make_partition {1e3*(t2-t1):.1f} ms, run(20) {1e3*(t3-t2):.1f} ms, close {1e3*(t4-t3):.1f} ms", file=sys.stderr)
