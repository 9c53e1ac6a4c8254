import os, sys, statistics, json
sys.path.insert(0, '/root/repo')
import torch
from voxelized_geometry_tools_b200 import device as vdev, sharded, synthetic
dev = torch.device('cuda', 0)
def timed(fn, reps=3):
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
dims=(1024,1024,1024)
occ = synthetic.clustered_spheres_occupancy_torch(dims, dev)
packed = vdev.edt_local_passes(occ)
slab = occ[0:128]
yslab = packed[:, 0:128, :].contiguous()
work = torch.empty_like(yslab)
full_out = torch.empty_like(occ); mm = torch.empty(2, device=dev)
for name, env in [("default", {}), ("budget100000", {"VGT_B200_WINDOW_BUDGET": "100000"}),
                  ("budget9600", {"VGT_B200_WINDOW_BUDGET": "9600"}),
                  ("nopilot", {"VGT_B200_WINDOW_PILOT": "0"}), ("lean", {"VGT_B200_ENVELOPE": "lean"})]:
    for k in ("VGT_B200_WINDOW_BUDGET","VGT_B200_WINDOW_PILOT","VGT_B200_ENVELOPE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    vdev.edt_local_passes(slab, send_parts=8)
    t_local = timed(lambda: vdev.edt_local_passes(slab, send_parts=8))
    t_local_plain = timed(lambda: vdev.edt_local_passes(slab))
    def final():
        work.copy_(yslab); vdev.edt_final_pass(work, 0, 1024, 0.02)
    final()
    t_copy = timed(lambda: work.copy_(yslab))
    t_final = timed(final) - t_copy
    whole = vdev.signed_distance_field_profile(occ, 0.02, full_out, mm)
    whole = vdev.signed_distance_field_profile(occ, 0.02, full_out, mm)
    print(json.dumps({"case": name, "slab_local_send": t_local, "slab_local_plain": t_local_plain, "slab_final": t_final, "whole": whole}), flush=True)
