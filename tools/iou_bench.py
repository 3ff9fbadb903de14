"""IoU kernel alone: rvb_iou_batch_run on device-resident boxes, kernel time from the library's CUDA events.
usage: python tools/iou_bench.py [frames_per_video] [videos] [annotators]   (RVB_IOU_GENERIC=1: the run-time annotator loop)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from retargetvid_b200 import _cabi

fpv = int(sys.argv[1]) if len(sys.argv) > 1 else 613
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
U = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ctx = _cabi.Context(0)
foff = np.arange(nv + 1, dtype=np.int64) * fpv
nev = np.full(nv, fpv, dtype=np.int32)
nfi = int(foff[-1])
g = torch.Generator(device='cuda').manual_seed(7)
mx = torch.randint(0, 500, (nfi, 1), device='cuda', generator=g, dtype=torch.int32)
z = torch.zeros((nfi, 1), device='cuda', dtype=torch.int32)
method = torch.cat([mx, z, mx + 119, z + 359], dim=1).contiguous()
x1 = torch.randint(0, 500, (U, nfi, 1), device='cuda', generator=g, dtype=torch.int32)
y1 = torch.zeros((U, nfi, 1), device='cuda', dtype=torch.int32)
annot = torch.cat([x1, y1, x1 + 120, y1 + 360], dim=2).contiguous()
acc = torch.zeros((nv, U, 2), dtype=torch.int64, device='cuda')
ib = _cabi.rvb_iou_batch()
ib.n_videos, ib.n_users, ib.mem_space = nv, U, _cabi.RVB_MEM_DEVICE
ib.frame_offset, ib.n_eval = foff.ctypes.data, nev.ctypes.data
ib.method_boxes, ib.annot_boxes, ib.frame_iou, ib.acc = method.data_ptr(), annot.data_ptr(), None, acc.data_ptr()
for _ in range(3):
	ctx.iou_batch(ib)
ms = []
for _ in range(20):
	ctx.iou_batch(ib)
	ms.append(ctx.last_iou_kernel_ms())
ms = float(np.median(ms))
by = nfi * (16 + 16 * U)
print(json.dumps({'frames_per_video': fpv, 'videos': nv, 'annotators': U, 'generic': bool(os.environ.get('RVB_IOU_GENERIC')),
				'kernel_ms': ms, 'gbs': by / ms / 1e6, 'acc_checksum': int(acc.sum().item())}))
