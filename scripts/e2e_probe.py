"""diagnostic: where does the host-buffer call (tlc_vicinity_pi) spend its time? (run on the GPU box)
   wall time per step of (a) the device-buffer call, (b) the C-ABI host-buffer call with a reused output array, (c) the
   same with a fresh output array per call, (d) the sg2dgm mirror -- next to the device time of the stage timers"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tlc-gnn_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
os.environ["TLC_STAGE_TIMING"] = os.environ.get("TLC_STAGE_TIMING", "1")
import bench
from tlc_b200 import api, _lib as L
import sg2dgm.riccidist2dgm as mirror

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
c, labels, ne, csr, perm = bench.make_workload("computers")
gm = mirror.graph2pi.from_csr(*csr, device=0)
g = gm._graph
batches = [bench.batch_targets(ne, perm, 500 + s, 0, 1, B) for s in range(8)]
dev = torch.device("cuda", 0)
out = np.zeros((B, 25))
d_t = [torch.from_numpy(b).to(dev) for b in batches]
d_pi = torch.empty((B, 25), dtype=torch.float64, device=dev)
d_st = torch.empty((B,), dtype=torch.uint8, device=dev)

def run(name, fn):
    fn(0); fn(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for s in range(2, 8):
        fn(s)
        dev_ms += g.last_stage_ms()[0].get("total", 0.0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 6
    print("%-34s wall %.3f ms/step   device (stage timers) %.3f ms/step" % (name, dt * 1e3, dev_ms / 6), flush=True)

run("device buffers", lambda s: (g.vicinity_pi_dev(d_t[s], d_pi, out_status=d_st, hop=2, flags=L.F_NORM), torch.cuda.synchronize()))
run("host buffers, reused output", lambda s: g.vicinity_pi(batches[s], hop=2, flags=L.F_NORM, out=out))
run("host buffers, fresh output", lambda s: g.vicinity_pi(batches[s], hop=2, flags=L.F_NORM))
run("sg2dgm mirror", lambda s: gm.get_pimg_for_all_edges(batches[s], cores=16, hop=2, norm=True, extended_flag=False, resolution=5, descriptor="sum"))
