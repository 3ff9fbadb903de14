import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from retargetvid_b200 import smartVidCrop as svc, synth
from retargetvid_b200.engine import CropEngine
e = CropEngine(0)
vds = [synth.make_clip(100+i, fc=40, shot_starts=[18] if i else []) for i in range(2)]
for best in (False, True):
    CP = svc.sc_init_crop_params(use_best_settings=best)
    r = e.run(vds, CP, ['1:3','3:1'], detail=True, want_filtered=True)
    print('ok', best, r[0].boxes[0][0], r[1].status)
CP = svc.sc_init_crop_params(); CP['clust_filt']=False
print(e.run(vds, CP, ['1:3'])[0].boxes[0][0])
