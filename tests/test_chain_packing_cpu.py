"""Replays, on the CPU, what conv_chain_kernel is handed for a fused chain (esr_debug_chain: per layer the three dy
weight parts with the dx taps interleaved atom by atom, the centre block, the bias table and the layer record) in the
kernel's own formulation - row-stationary stacked MMAs: input row i times [dy = +1 | 0 | -1] parts lands in output
rows i-2, i-1, i of three consecutive accumulator regions - and compares every layer with the oracle's evaluation of
the reference layers folded into it.  Pins the host side of the fused kernel (find_chains / chain_pack_layer) without
a GPU.  Weights are compared at their fp16-rounded values (what the blob holds); arithmetic is fp64, and the
activations are rounded to fp16 between the layers exactly where the kernel's epilogue does."""
import ctypes
import os

import numpy as np
import pytest

from oracle import esr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTR_BYTES = 4096


def _weights(mid):
    return O.load_weights(os.path.join(ROOT, "tests", "golden", "weights", O.MODELS[mid]["weights"] + ".npz"))


def _h(a):   # double -> float -> half, as the packer rounds
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float16).astype(np.float64)


def _f16(a):
    return np.asarray(a, dtype=np.float64).astype(np.float16).astype(np.float64)


def _lrelu(x, s=0.05):
    return np.where(x >= 0, x, s * x)


def _names(arch, w, **kw):
    """tcgen05 layer index -> op name, through the per-layer export"""
    from ntire2022_esr_b200 import Engine, _cabi
    e = Engine(arch, device=-1, **kw)
    e.load_state_dict(w)
    lib = _cabi.lib
    n = lib.esr_debug_tc_layer(e._h, -1, None, 0, None, None, None, None, None, None, 0)
    names = []
    for i in range(n):
        name = ctypes.create_string_buffer(256)
        meta = (ctypes.c_int32 * 12)(); ent = (ctypes.c_int32 * 128)(); grp = (ctypes.c_int32 * 24)()
        bias = (ctypes.c_float * 192)(); bias9 = (ctypes.c_float * 576)()
        assert lib.esr_debug_tc_layer(e._h, i, name, 256, meta, ent, grp, bias, bias9, None, 0) == 0
        names.append(name.value.decode())
    return e, names


def _chains(arch, w, **kw):
    from ntire2022_esr_b200 import _cabi
    e, names = _names(arch, w, **kw)
    lib = _cabi.lib
    n = lib.esr_debug_chain(e._h, -1, None, None, None, 0)
    out = []
    for i in range(n):
        meta = (ctypes.c_int32 * 4)()
        lay = (ctypes.c_int32 * (12 * 8))()
        blob = (ctypes.c_uint8 * (1 << 20))()
        assert lib.esr_debug_chain(e._h, i, meta, lay, blob, len(blob)) == 0, lib.esr_last_error(e._h)
        L = np.array(lay, dtype=np.int64).reshape(8, 12)[:meta[0]]
        out.append(dict(layers=[dict(name=names[r[0]], np=int(r[1]), ksteps=int(r[2]), ctr_n=int(r[3]), part_bytes=int(r[4]), off=int(r[5]),
                                     ident=int(r[6]), n0=int(r[7]), n1=int(r[8]), g1_ctr=int(r[9]), col1=int(r[10])) for r in L],
                        blob=np.frombuffer(bytes(blob[:meta[1]]), dtype=np.uint8), pw=bool(meta[3])))
    return out


def _half_at(blob, off):
    return (blob[off].astype(np.uint16) | (blob[off + 1].astype(np.uint16) << 8)).view(np.float16).astype(np.float64)


def _unpack(blob, L):
    """-> parts[q][n][dxi][k] (q: dy = +1, 0, -1), ctr[n][k], bias0[64], bias1[64]"""
    npad = L["np"]
    n, d, k = np.meshgrid(np.arange(npad), np.arange(3), np.arange(64), indexing="ij")
    a, r = n >> 3, n & 7
    rel = (a * 3 + d) * 1024 + r * 128 + ((((k >> 3) ^ r) & 7) << 4) + (k & 7) * 2           # chain_host.cuh::chain_pack_layer
    parts = np.stack([_half_at(blob, L["off"] + q * L["part_bytes"] + rel) for q in range(3)])
    assert L["part_bytes"] == npad * 384
    n, k = np.meshgrid(np.arange(32), np.arange(64), indexing="ij")
    sw = (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 3) ^ (n & 7)) & 7) << 4) + (k & 7) * 2  # tc_common.cuh::sw128_offset
    ctr = _half_at(blob, L["off"] + 3 * L["part_bytes"] + sw)
    bias = np.frombuffer(blob[L["off"] + 3 * L["part_bytes"] + CTR_BYTES:][:512].tobytes(), dtype=np.float32).astype(np.float64)
    return parts, ctr, bias[:64], bias[64:]


def _replay_layer(blob, L, x):
    """x (H, W, 64) -> accumulator (H, W, np) and centre accumulator (H, W, 32), the way the kernel forms them"""
    H, W, _ = x.shape
    parts, ctr, b0, b1 = _unpack(blob, L)
    K = 16 * L["ksteps"]
    assert not parts[..., K:].any() and not ctr[:, K:].any(), "weights beyond the issued K steps must be zero"
    assert not ctr[L["ctr_n"]:].any()
    xp = np.zeros((H + 2, W + 2, 64))
    xp[1:-1, 1:-1] = x
    acc = np.zeros((H, W, L["np"]))
    for yi in range(-1, H + 1):                       # input row (image coordinates; rows -1 and H are zero padding)
        row = xp[yi + 1]
        for dxi in range(3):
            a = row[dxi:dxi + W, :K]                  # the dx-shifted view of the row
            stacked = np.concatenate([parts[q][:, dxi, :K] for q in range(3)])     # N = 3 np: [dy = +1 | 0 | -1]
            out = a @ stacked.T                        # one MMA per K step on the hardware
            for q in range(3):
                yo = yi - (1 - q)                      # output row of part q
                if 0 <= yo < H:
                    acc[yo] += out[:, q * L["np"]:(q + 1) * L["np"]]
    if L["ident"]:
        acc[..., :64] += x[..., :L["np"]] if L["np"] < 64 else x
    cacc = x[..., :K] @ ctr[:, :K].T
    return acc, cacc, b0, b1


def test_rfdn_block_chain_against_the_reference_layers():
    """conv_chain:B2.c1_r+d | c2_r+d | c3_r+d | c4 (block.py:148-163): r_k = lrelu(c_k_r(r_{k-1}) + r_{k-1}),
    d_k = lrelu(c_k_d(r_{k-1})), r4 = lrelu(c4(r3))."""
    w = _weights(0)
    ch = [c for c in _chains("rfdn", w) if c["layers"][0]["name"].startswith("B2.c1_r")][0]
    assert [l["name"] for l in ch["layers"]] == ["B2.c1_r+d", "B2.c2_r+d", "B2.c3_r+d", "B2.c4"]
    rng = np.random.default_rng(0)
    x = np.zeros((7, 11, 64))
    x[..., :50] = _f16(rng.standard_normal((7, 11, 50)))
    for k, L in enumerate(ch["layers"]):
        acc, cacc, b0, b1 = _replay_layer(ch["blob"], L, x)
        xin = x[..., :50].transpose(2, 0, 1)[None]
        if k < 3:
            assert (L["np"], L["ksteps"], L["ctr_n"], L["ident"], L["n0"], L["n1"], L["g1_ctr"]) == (64, 4, 32, 1, 64, 32, 1)
            r = O.conv2d(xin, _h(w[f"B2.c{k + 1}_r.weight"]), None, 1, 1)[0].transpose(1, 2, 0) + x[..., :50]
            d = O.conv2d(xin, _h(w[f"B2.c{k + 1}_d.weight"]), None)[0].transpose(1, 2, 0)
            assert np.abs(acc[..., :50] - r).max() < 1e-10 and not acc[..., 50:].any()
            assert np.abs(cacc[..., :25] - d).max() < 1e-10 and not cacc[..., 25:].any()
            np.testing.assert_array_equal(b0[:50], w[f"B2.c{k + 1}_r.bias"].astype(np.float64))
            np.testing.assert_array_equal(b1[:25], w[f"B2.c{k + 1}_d.bias"].astype(np.float64))
            assert not b0[50:].any() and not b1[25:].any()
            nxt = np.zeros_like(x)
            nxt[..., :50] = _f16(_lrelu(acc[..., :50] + b0[:50]))      # the epilogue writes the next layer's input as fp16
            x = nxt
        else:
            assert (L["np"], L["ctr_n"], L["ident"], L["n1"]) == (32, 0, 0, 0)
            r4 = O.conv2d(xin, _h(w["B2.c4.weight"]), None, 1, 1)[0].transpose(1, 2, 0)
            assert np.abs(acc[..., :25] - r4).max() < 1e-10 and not acc[..., 25:].any()
            np.testing.assert_array_equal(b0[:25], w["B2.c4.bias"].astype(np.float64))


def test_tail_chain_lr_conv_then_upsampler():
    """conv_chain:LR_conv | upsampler (RFDN.py:36-41): out_lr = LR_conv(out_B) + out_fea is formed by the first layer
    (the residual `+ out_fea` is that layer's epilogue residual or identity input, checked through the oracle), the
    second layer is the 3x3 conv in front of PixelShuffle(4): 48 output columns."""
    w = _weights(0)
    ch = _chains("rfdn", w)[-1]
    assert [l["name"] for l in ch["layers"]] == ["LR_conv", "upsampler"] and not ch["pw"]
    rng = np.random.default_rng(1)
    x = np.zeros((6, 9, 64))
    x[..., :50] = _f16(rng.standard_normal((6, 9, 50)))
    L0, L1 = ch["layers"]
    acc, _, b0, _ = _replay_layer(ch["blob"], L0, x)
    ref = O.conv2d(x[..., :50].transpose(2, 0, 1)[None], _h(w["LR_conv.weight"]), None, 1, 1)[0].transpose(1, 2, 0)
    assert np.abs(acc[..., :50] - ref).max() < 1e-10 and L0["ctr_n"] == 0 and L0["n1"] == 0
    np.testing.assert_array_equal(b0[:50], w["LR_conv.bias"].astype(np.float64))
    acc1, _, bu, _ = _replay_layer(ch["blob"], L1, x)
    ref1 = O.conv2d(x[..., :50].transpose(2, 0, 1)[None], _h(w["upsampler.0.weight"]), None, 1, 1)[0].transpose(1, 2, 0)
    assert L1["np"] == 48 and np.abs(acc1[..., :48] - ref1).max() < 1e-10
    np.testing.assert_array_equal(bu[:48], w["upsampler.0.bias"].astype(np.float64))


def test_imdn_block_chain_channel_split():
    """conv_chain:model.1.sub.k.conv1..conv4 (basicblock.py:259-265): each conv's output columns are permuted to
    [remaining 48 | distilled 16], group 0 (48 columns) feeds the next layer, group 1 (16 columns of the same
    accumulator, not a centre block) is the distilled slice."""
    w = _weights(-1)
    ch = [c for c in _chains("imdn", w) if c["layers"][0]["name"].startswith("model.1.sub.1.conv1")][0]
    names = [l["name"] for l in ch["layers"]]
    assert len(names) == 4 and names[0].startswith("model.1.sub.1.conv1") and names[3].startswith("model.1.sub.1.conv4")
    rng = np.random.default_rng(2)
    x = _f16(rng.standard_normal((5, 8, 64)))
    cin = 64
    for k, L in enumerate(ch["layers"][:3]):
        assert (L["np"], L["ctr_n"], L["ident"], L["n0"], L["n1"], L["g1_ctr"], L["col1"]) == (64, 0, 0, 48, 16, 0, 48)
        acc, _, b0, b1 = _replay_layer(ch["blob"], L, x)
        wk = _h(w[f"model.1.sub.1.conv{k + 1}.0.weight"]) if f"model.1.sub.1.conv{k + 1}.0.weight" in w else _h(w[f"model.1.sub.1.conv{k + 1}.weight"])
        bk = w.get(f"model.1.sub.1.conv{k + 1}.0.bias", w.get(f"model.1.sub.1.conv{k + 1}.bias")).astype(np.float64)
        full = O.conv2d(x[..., :cin].transpose(2, 0, 1)[None], wk, None, 1, 1)[0].transpose(1, 2, 0)     # (H, W, 64): [distilled 16 | remaining 48]
        assert np.abs(acc[..., :48] - full[..., 16:]).max() < 1e-10 and np.abs(acc[..., 48:64] - full[..., :16]).max() < 1e-10
        np.testing.assert_array_equal(b0[:48], bk[16:])
        np.testing.assert_array_equal(b1[:16], bk[:16])
        nxt = np.zeros_like(x)
        nxt[..., :48] = _f16(_lrelu(acc[..., :48] + b0[:48]))
        x, cin = nxt, 48


@pytest.mark.parametrize("arch,mid,kw,nchains", [("rfdn", 0, {}, 5), ("imdn", -1, {}, 9), ("rlfn", 4, {}, 5), ("rfdn", 22, {"nf": 40}, 5),
                                                 ("rfdn_pruned", 40, {}, None), ("bsrn", 18, {}, 0)])
def test_chain_records_are_consistent(arch, mid, kw, nchains):
    """Structural invariants of every chain of every network: widths, K steps, part sizes, blob extents, and that BSRN
    (border-class bias) is never chained."""
    chains = _chains(arch, _weights(mid), **kw)
    if nchains is not None:
        assert len(chains) == nchains, [[l["name"] for l in c["layers"]] for c in chains]
    for c in chains:
        assert 2 <= len(c["layers"]) <= 6
        end = 0
        for L in c["layers"]:
            assert L["np"] in (16, 32, 48, 64) and 1 <= L["ksteps"] <= 4 and L["part_bytes"] == L["np"] * 384, L
            assert L["ctr_n"] in (0, 16, 32) and L["n0"] <= L["np"] and L["n1"] in (0, 16, 32), L
            assert L["off"] % 128 == 0 and L["off"] >= end, L
            end = L["off"] + 3 * L["part_bytes"] + CTR_BYTES + 512
            assert end <= len(c["blob"]) + 127, L
            if L["n1"] and not L["g1_ctr"]:
                assert L["col1"] >= L["n0"] and L["col1"] + L["n1"] <= L["np"], L
        assert c["layers"][-1]["n1"] == 0            # the last layer of a chain stores group 0 only
