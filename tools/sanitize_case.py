"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel of the library once,
including the last-rows-of-the-pool staging fallback and the generic crop path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpc_baseline_b200 import batched, synth

dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
batch = synth.make_scenes(6, 12, seed=5, p_drop=0.2, n_dup=1, n_false=1, width=800, height=600, side_lo=16, side_hi=300)
Ks, RTs, cen, cnt, boxes = dev(batch.Ks), dev(batch.RTs), dev(batch.centers), dev(batch.counts), dev(batch.boxes)
res = batched.match_triangulate(Ks, RTs, cen, cnt, 30, want_F=True)
F = batched.fundamental(Ks, RTs)
cost = batched.cost_tensor(F, cen, cnt)
batched.match_objects(cost[:, :4, :4, :4].contiguous(), 30)
batched.box_centers(boxes)
images = dev(synth.make_images(3, seed=6, width=800, height=600))
ios = dev(np.tile(np.arange(3, dtype=np.int32), (6, 1)))
rois, offs = batched.build_rois(boxes, res.idx, res.n, ios)
n = int(offs[-1])
extra = dev(np.array([[2, 0, 0, 800, 600], [2, 500, 380, 800, 600], [2, 790, 0, 800, 600], [2, 0, 590, 800, 600],
                      [0, 10, 10, 458, 458], [1, 5, 5, 229, 229], [1, 100, 100, 130, 140]], np.int32))
allrois = torch.cat([rois[:n], extra]).contiguous()
for T in (224, 256, 50):
    batched.roi_crop(images, allrois, T=T)
    batched.roi_crop_u8(images, allrois, T=T)
P = batched.projection(Ks.reshape(-1, 3, 3), RTs.reshape(-1, 4, 4))
torch.cuda.synchronize()
print('sanitize case ok', n, 'rois +', len(extra))
