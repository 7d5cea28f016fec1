"""CPU-only checks: the C-ABI library loads and exports every symbol include/esr_b200.h declares,
host-side weight checking (strict load_state_dict semantics), and the drop-in module facade.
No compute call is made here (there is no CPU path in the engine)."""
import os
import re
import types

import numpy as np
import pytest
import torch

from oracle import esr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHS = [(-1, "imdn"), (0, "rfdn"), (4, "rlfn"), (18, "bsrn")]


def _weights(mid):
    return O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))


def test_library_exports_every_declared_symbol():
    from ntire2022_esr_b200 import _cabi

    header = open(os.path.join(ROOT, "include", "esr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(esr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    bound = {s[0] for s in _cabi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(_cabi.lib, name)
    assert b"sm_100a" in _cabi.lib.esr_version()
    assert _cabi.lib.esr_device_ok(-1) == 0


@pytest.mark.parametrize("mid,arch", ARCHS)
def test_strict_state_dict_and_plan_names(mid, arch):
    from ntire2022_esr_b200 import Engine, EsrError, _cabi

    w = _weights(mid)
    e = Engine(arch, device=-1)
    e.load_state_dict(w)
    n32 = e.launch_names(1, 64, 64, _cabi.DTYPE_F32)
    n16 = e.launch_names(1, 64, 64, _cabi.DTYPE_F16)
    tc = ("conv_tc", "conv_chain")      # the two tcgen05 kernels: per-layer and fused chain
    assert n32 and n16 and not any(n.startswith(tc) for n in n32)
    assert any(n.startswith(tc) for n in n16)
    if arch != "bsrn":   # the stacked 3x3 layers of a block run as one fused launch (BSRN's border-class bias keeps it per-layer)
        assert any(n.startswith("conv_chain") for n in n16)
    e.set_option("chain_enable", 0)
    assert not any(n.startswith("conv_chain") for n in e.launch_names(1, 64, 64, _cabi.DTYPE_F16))
    e.set_option("chain_enable", 1)
    assert e.workspace_bytes(2, 64, 48, _cabi.DTYPE_F16) > 0
    with pytest.raises(EsrError) as ei:  # no GPU bound -> loud failure, never a CPU fallback
        e.forward_host(np.zeros((1, 3, 32, 32), np.float32))
    assert ei.value.code == _cabi.E_NOGPU
    if arch != "imdn":
        with pytest.raises(EsrError):  # ESA needs H, W >= 15
            e.workspace_bytes(1, 14, 20, _cabi.DTYPE_F32)
    # missing key
    k = sorted(w)[5]
    e2 = Engine(arch, device=-1)
    with pytest.raises(EsrError, match="missing key"):
        e2.load_state_dict({a: b for a, b in w.items() if a != k})
    # unexpected key
    e3 = Engine(arch, device=-1)
    with pytest.raises(EsrError, match="unexpected key"):
        e3.load_state_dict({**w, "extra.weight": np.zeros(3, np.float32)})
    # wrong shape
    e4 = Engine(arch, device=-1)
    bad = dict(w)
    bad[k] = np.zeros(tuple(s + 1 for s in w[k].shape), np.float32)
    with pytest.raises(EsrError, match="size mismatch"):
        e4.load_state_dict(bad)


@pytest.mark.parametrize("mid,arch", ARCHS)
def test_module_facade_matches_reference_state_dict_contract(mid, arch):
    from ntire2022_esr_b200 import EsrError, build_model

    w = _weights(mid)
    m = build_model(mid, state_dict=w)
    sd = m.state_dict()
    assert list(sd.keys()) == list(w.keys())          # same names, same order as the reference module
    for k in w:
        assert tuple(sd[k].shape) == w[k].shape
        np.testing.assert_array_equal(sd[k].numpy(), w[k])
    assert isinstance(m, torch.nn.Module)
    assert sum(p.numel() for p in m.parameters()) == sum(v.size for v in w.values())
    m.eval()
    assert len(list(m.modules())) > 1
    with pytest.raises(RuntimeError):                  # strict=True
        m.load_state_dict({k: torch.as_tensor(v) for k, v in list(w.items())[:-1]}, strict=True)
    with pytest.raises(EsrError):                      # CPU tensors are refused, not silently computed
        m(torch.zeros(1, 3, 32, 32))


def test_select_model_contract(tmp_path, monkeypatch):
    from ntire2022_esr_b200 import select_model
    from ntire2022_esr_b200 import specs

    zoo = tmp_path / "model_zoo"
    zoo.mkdir()
    for mid, arch in ARCHS:
        reg = specs.REGISTRY[mid]
        sd = {k: torch.from_numpy(v) for k, v in _weights(mid).items()}
        torch.save({reg["wrap"]: sd} if reg["wrap"] else sd, zoo / reg["file"])
    monkeypatch.chdir(tmp_path)                         # reference paths are CWD-relative
    names = {-1: "-1_IMDN_baseline", 0: "00_RFDN_baseline", 4: "04_RLFN", 18: "18_RFDNFINALB5"}
    for mid, arch in ARCHS:
        model, name, data_range, tile = select_model(types.SimpleNamespace(model_id=mid), torch.device("cpu"))
        assert name == names[mid] == O.MODELS[mid]["name"]
        assert data_range == O.MODELS[mid]["data_range"]
        assert tile is None
        assert not model.training and all(not p.requires_grad for p in model.parameters())
    with pytest.raises(NotImplementedError):
        select_model(types.SimpleNamespace(model_id=99), torch.device("cpu"))


def test_tiled_forward_accumulate_and_divide():
    """forward(tile=...) bookkeeping (test_demo.py:368-389) with a stand-in model on CPU."""
    from ntire2022_esr_b200 import forward

    class Up(torch.nn.Module):
        def forward(self, x):
            y = torch.nn.functional.interpolate(x, scale_factor=4, mode="nearest")
            return y + x.mean()                         # depends on the tile content

    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 37, 29, generator=g)
    out = forward(x, Up(), tile=16, tile_overlap=4)
    # independent restatement
    tile, stride = 16, 12
    ys = list(range(0, 37 - tile, stride)) + [37 - tile]
    xs = list(range(0, 29 - tile, stride)) + [29 - tile]
    e = np.zeros((2, 3, 148, 116), np.float64)
    wsum = np.zeros_like(e)
    for a in ys:
        for b in xs:
            p = x[..., a:a + tile, b:b + tile]
            o = Up()(p).double().numpy()
            e[..., 4 * a:4 * (a + tile), 4 * b:4 * (b + tile)] += o
            wsum[..., 4 * a:4 * (a + tile), 4 * b:4 * (b + tile)] += 1
    np.testing.assert_allclose(out.numpy(), e / wsum, rtol=1e-6, atol=1e-6)
    assert out.shape == (2, 3, 148, 116)
    torch.testing.assert_close(forward(x, Up(), tile=None), Up()(x))


@pytest.mark.parametrize("mid", [22, 40])
def test_next_row_models_load_strictly_and_plan(mid):
    """RFDN40 (id 22) and the pruned RFDN (id 40) reuse the RFDN graph builder; the module facade
    takes the reference state-dict with strict=True and the engine plans both flavours (host-only handle)."""
    from ntire2022_esr_b200 import Engine, _cabi, build_model, specs

    w = _weights(mid)
    reg = specs.REGISTRY[mid]
    m = build_model(mid, state_dict=w)
    assert set(m.state_dict()) == set(w)
    e = Engine(reg["arch"], device=-1, nf=reg["kwargs"]["nf"], nblocks=reg["kwargs"]["nblocks"])
    e.load_state_dict(w)
    n16 = e.launch_names(1, 64, 64, _cabi.DTYPE_F16)
    # fused: one chain per RFDB (c1_r+d .. c4) and one for LR_conv + upsampler; c5 (+ ESA entry) and `c` stay plain launches
    assert sum(n.startswith("conv_chain") for n in n16) == 5 and sum(n.startswith("conv_tc") for n in n16) == 5
    assert len(n16) == 1 + 1 + 4 * 5 + 1 + 1            # flag memset, head, 4 x (chain + c5 + 3 ESA launches), c, tail chain
    e.set_option("chain_pw", 1)                          # option: c5 as the chain's last (pointwise) stage
    n16 = e.launch_names(1, 64, 64, _cabi.DTYPE_F16)
    assert sum(n.startswith("conv_chain") for n in n16) == 5 and sum(n.startswith("conv_tc") for n in n16) == 1
    e.set_option("chain_pw", 0)
    e.set_option("chain_enable", 0)
    n16 = e.launch_names(1, 64, 64, _cabi.DTYPE_F16)
    assert sum(n.startswith("conv_tc") for n in n16) == 23 and len(n16) == 36
    with pytest.raises(Exception, match="size mismatch|missing key|unexpected key"):
        Engine("rfdn", device=-1, nf=50).load_state_dict(w)      # a 40-channel checkpoint in the nf = 50 graph


def test_uint8_entry_points_refuse_without_a_gpu():
    """esr_forward_u8 / esr_forward_host_u8 on a host-only handle: loud ESR_E_NOGPU, never a CPU fallback; the
    workspace query works without a device (it is pure arithmetic on the plan)."""
    from ntire2022_esr_b200 import Engine, EsrError, _cabi

    e = Engine("rfdn", device=-1)
    e.load_state_dict(_weights(0))
    base = _cabi.lib.esr_workspace_bytes(e._h, 2, 64, 48, _cabi.DTYPE_F16)
    u8 = _cabi.lib.esr_workspace_bytes_u8(e._h, 2, 64, 48, _cabi.DTYPE_F16)
    assert u8 >= base + 2 * 3 * 16 * 64 * 48 * 2           # + the engine's own NCHW output
    with pytest.raises(EsrError) as ei:
        e.forward_host_uint8(np.zeros((1, 32, 32, 3), np.uint8), 255.0)
    assert ei.value.code == _cabi.E_NOGPU
    with pytest.raises(EsrError):
        e.forward_host_uint8(np.zeros((1, 32, 32, 3), np.float32), 255.0)   # wrong dtype


# utils/model_summary.py numbers of the UNMODIFIED reference for input (3, 256, 256), produced in the build container
# with get_model_activation / get_model_flops on test_demo.select_model(id) (test_demo.py:522-533):
# id -> (#activations, #Conv2d, FLOPs, #params)
REFERENCE_SUMMARY = {
    -1: (154140672, 43, 58531512320, 893936), 0: (112034240, 64, 27102622560, 433448), 4: (80045184, 39, 19695306752, 317218),
    18: (65757744, 43, 2022422883, 156288), 22: (90500128, 64, 17613787280, 282688), 26: (136314880, 38, 51757711360, 790496),
    40: (91718080, 64, 17700828000, 289888), 3: (72089600, 34, 22280011776, 341066),
}


@pytest.mark.parametrize("mid", sorted(REFERENCE_SUMMARY))
def test_model_summary_hooks_see_the_reference_layers(mid, monkeypatch):
    """The harness prints #Activations / #Conv2d / FLOPs / #Params of the selected model through forward hooks on its
    nn.Conv2d / nn.Linear / nn.ReLU / nn.LeakyReLU leaves (test_demo.py:522-533, utils/model_summary.py:230-245,390-400).
    The drop-in keeps its weights in real leaves of the reference's types and serves the hooks with shape-only tensors
    after the engine forward, so the same code reports the reference's numbers.  The hook arithmetic is restated here
    (model_summary.py:283-318, 428-438); with /root/reference present the reference's own functions are used as well."""
    import numpy as np
    import torch.nn as nn

    from ntire2022_esr_b200 import build_model

    m = build_model(mid, state_dict=_weights(mid))

    class _NoEngine:   # the CPU suite has no GPU: stand-in for the engine call (the hooks never look at values)
        def forward(self, x):
            return torch.zeros(x.shape[0], 3, 4 * x.shape[2], 4 * x.shape[3])

    monkeypatch.setattr(m, "engine", lambda dev: _NoEngine())
    counts = {"flops": 0, "act": 0, "nconv": 0}

    def conv_hook(mod, inp, out):
        counts["flops"] += int(np.prod(mod.kernel_size) * mod.in_channels * (mod.out_channels // mod.groups)) * int(out.shape[0] * np.prod(out.shape[2:]))
        counts["act"] += out.numel()
        counts["nconv"] += 1

    def relu_hook(mod, inp, out):
        counts["flops"] += out.numel()

    def linear_hook(mod, inp, out):
        counts["flops"] += int(inp[0].shape[0] * inp[0].shape[1] * out.shape[1])

    handles = []
    for mod in m.modules():
        if isinstance(mod, nn.Conv2d):
            handles.append(mod.register_forward_hook(conv_hook))
        elif isinstance(mod, (nn.ReLU, nn.LeakyReLU)):
            handles.append(mod.register_forward_hook(relu_hook))
        elif isinstance(mod, nn.Linear):
            handles.append(mod.register_forward_hook(linear_hook))
    m(torch.zeros(1, 3, 256, 256))
    for h in handles:
        h.remove()
    want = REFERENCE_SUMMARY[mid]
    assert (counts["act"], counts["nconv"], counts["flops"], sum(p.numel() for p in m.parameters())) == want
    if os.path.isdir("/root/reference/utils"):
        import sys
        import types

        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.path.insert(0, "/root/reference")
        try:
            from utils.model_summary import get_model_activation, get_model_flops
        finally:
            sys.path.remove("/root/reference")
        assert get_model_activation(m, (3, 256, 256)) == want[:2]
        assert get_model_flops(m, (3, 256, 256), False) == want[2]


def test_fmen_loads_strictly_and_the_chain_planner_keeps_the_gate_operands_in_global_memory():
    """SURVEY row N1, FMEN (id 3, models/team03_fmen.py:78-134): 34 3x3 convolutions.  An HFAB multiplies its own input
    with the sigmoid of its last convolution, so the buffer holding that input must be complete in global memory when
    the gate layer's epilogue reads it: the planner ends a fused chain at the layer that produces it, never lets a chain
    write a buffer one of its own layers still reads, and plans at most four layers per launch."""
    from ntire2022_esr_b200 import Engine, EsrError, _cabi, build_model

    w = _weights(3)
    m = build_model(3, state_dict=w)
    assert set(m.state_dict()) == set(w) and sum(p.numel() for p in m.parameters()) == 341066
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in m.state_dict().items() if k != "lr_conv.bias"}, strict=True)
    e = Engine("fmen", device=-1)
    e.load_state_dict(w)
    n16 = e.launch_names(1, 64, 64, _cabi.DTYPE_F16)
    chains = [n[len("conv_chain:"):].split(" | ") for n in n16 if n.startswith("conv_chain")]
    assert n16[1] == "head:head" and n16[2] == "conv_tc:warmup.0"           # warm-up conv: its output is HFAB 0's gate
    assert all(2 <= len(c) <= 4 for c in chains) and sum(len(c) for c in chains) + 1 == 33   # every conv but the head
    assert chains[-1] == ["lr_conv", "tail.0"]
    for c in chains:                                                      # a gate producer is always the last layer of its launch
        for name in c[:-1]:
            assert not (name.startswith("basic_blocks.") and name.endswith("conv2.rep_conv")), c
    assert len(e.launch_names(1, 64, 64, _cabi.DTYPE_F32)) == 34
    e.workspace_bytes(1, 5, 7, _cabi.DTYPE_F16)                          # no ESA: no 15-pixel minimum
    with pytest.raises(EsrError):
        Engine("fmen", device=-1).load_state_dict({k: v for k, v in w.items() if k != "head.weight"})
