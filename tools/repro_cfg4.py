import sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from groundgrid_b200 import capi, synth
which = sys.argv[1] if len(sys.argv) > 1 else "4x64"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
scan = {"4x64": synth.scan_4lidar, "64": synth.scan_64}[which]
pcap = 524288 if which == "4x64" else 131072
scene = synth.make_scene(seed=1, stream_len=2.0)
clouds = [scan(scene, (float(s), 0.0), 0.0, seed=s) for s in range(2)]
raw = [np.ascontiguousarray(c[0]).view(np.uint8).reshape(-1) for c in clouds]
pool = torch.from_numpy(np.concatenate(raw)).cuda()
offs = [0, raw[0].size]
g = capi.GroundGridB200(120.0, 0.33, n_slots=B, max_points=pcap, full_layers=False)
for b in range(B):
    g.init_map(0.0, 0.0, 0.0, slot=b)
slots = np.arange(B, dtype=np.int32)
for t in range(4):
    s = t & 1
    if t:
        q, tt = synth.base_from_map_qt(float(s), 0.0)
        T = synth.tf2_matrix(q, tt)
        g.update_pose_batch(slots, np.tile(np.array([float(s), 0.0]), (B, 1)), np.tile(T.reshape(1, 12), (B, 1)))
    d = g.make_descs(list(range(B)), [len(clouds[s][0])] * B, [clouds[s][1]] * B, [0.0] * B)
    g.run_scans_device(d, [pool.data_ptr() + offs[s]] * B)
    g.synchronize()
    print("step", t, "ok", len(clouds[s][0]), flush=True)
