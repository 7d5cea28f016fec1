"""Developer probe (GPU box): repeated forwards of one shape with optional engine options, checking for launch
failures and bit-identical outputs.

    python tools/gpu_stress.py <rfdn|bsrn> B H W n_forwards sync_every use_graph [option=value ...]
"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from oracle import esr_oracle as O
from ntire2022_esr_b200 import Engine
arch, B, H, W, n, sync_every, graph = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
ids = {"rfdn": 0, "bsrn": 18}
w = O.load_weights(f"tests/golden/weights/{O.MODELS[ids[arch]]['weights']}.npz")
eng = Engine(arch, 0)
eng.set_option("use_graph", graph)
for kv in sys.argv[8:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
eng.load_state_dict(w)
x = (torch.rand(B, 3, H, W) * O.MODELS[ids[arch]]["data_range"]).half().cuda()
y = torch.empty(B, 3, 4 * H, 4 * W, dtype=torch.float16, device="cuda")
ref = None
for i in range(n):
    try:
        eng.forward(x, out=y)
        if (i + 1) % sync_every == 0:
            torch.cuda.synchronize()
    except Exception as e:
        print(f"{arch} B{B} {H}x{W} graph={graph} sync_every={sync_every} {sys.argv[8:]}: FAILED at forward {i}: {str(e)[:90]!r}", flush=True)
        sys.exit(0)
    if (i + 1) % sync_every == 0:
        if ref is None:
            ref = y.clone()
        elif not torch.equal(ref, y):
            print(f"forward {i}: output differs from the first one, max diff {(ref.float() - y.float()).abs().max().item()}", flush=True)
print(f"{arch} B{B} {H}x{W} graph={graph} sync_every={sync_every} {sys.argv[8:]}: {n} forwards ok", flush=True)
