import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import esr_oracle as O
from ntire2022_esr_b200 import build_model
for mid in (0, 22, 40, 4):
    w = O.load_weights(f'/root/repo/tests/golden/weights/{O.MODELS[mid]["weights"]}.npz')
    m = build_model(mid, state_dict=w).eval().to('cuda:0')
    eng = m.engine(torch.device('cuda:0'))
    dr = O.MODELS[mid]['data_range']
    for shape in ((1, 3, 256, 256), (2, 3, 33, 47), (1, 3, 130, 260)):
        x = (torch.rand(*shape, generator=torch.Generator().manual_seed(3)) * dr).half().cuda()
        outs, ts = [], []
        for old in (1, 0):
            eng.set_option('esa_front_old', old)
            y = eng.forward(x).clone()
            for _ in range(30): eng.forward(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(300): eng.forward(x)
            e1.record(); torch.cuda.synchronize()
            outs.append(y); ts.append(e0.elapsed_time(e1) / 300 * 1e3)
        print(O.MODELS[mid]['name'], shape, 'bit-equal' if torch.equal(outs[0], outs[1]) else 'DIFFERENT', f'old {ts[0]:.1f} us new {ts[1]:.1f} us', flush=True)
