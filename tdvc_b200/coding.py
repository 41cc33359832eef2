"""Real entropy coding of the two latents of a P-frame (`forward(..., is_compress=True)`).

Reference path: main/model/pnet.py:45-49,69-73 - `coder.eval(); coder.update(force=True); out_enc = coder.compress(x);
ac_bpp = sum(len(s[0]) for s in out_enc["strings"]) * 8.0 / num_pixels` for `mvCoder` and `resCoder` (compressai
`Cheng2020Anchor`; compressai is not in the reference tree: SURVEY.md App. A, DESIGN.md section 7, parity is defined by
oracle/compressai_port.py + oracle/rans.py).

What runs where:
* `update`: HOST code.  The tables are integer data that the two ends of a codec must reproduce bit for bit, so nothing
  in them goes through device transcendental functions: the pmf of the factorised prior (a 5-layer MLP on <= 128 x 60
  points) and of the 64 Gaussian scales is evaluated with the same fp32 host operations compressai uses for a model that
  lives on the CPU, and quantised to 16 bits by `tdvc_pmf_to_quantized_cdf` (host code in the library, as compressai's is).
  (Measured: with the pmf from expf / erfcf on the GPU a bin edge differs by one count in ~0.1 % of the bins, which then
  changes which bins the zero-probability tail steals from.)  Tables are cached per set of packed weights (the reference
  rebuilds them on every call).
* `compress`: z symbols by `tdvc_eb_symbols`; the autoregressive pass over y - compressai's per-position Python loop - by
  the wavefront kernel `tdvc_ar_code` (one launch per coder, csrc/coding.cu) on the y / hyper-decoder buffers the forward
  pass left in the plan; rANS over the symbols on the host (`tdvc_rans_encode_with_indexes`), as in compressai.
"""
import ctypes as C
import math
import statistics

import numpy as np
import torch

from tdvc_b200 import lib as L

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64   # compressai models/priors.py
TAIL_MASS = 1e-9
PRECISION = 16


def scale_table():
    """compressai `get_scale_table()` (host arithmetic there too: torch.linspace defaults to the CPU)."""
    return torch.exp(torch.linspace(math.log(SCALES_MIN), math.log(SCALES_MAX), SCALES_LEVELS))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Tables:
    """compressai's `_quantized_cdf`, `_cdf_length`, `_offset` of one entropy model, as int32 numpy arrays."""

    def __init__(self, cdf, length, offset):
        self.cdf = np.ascontiguousarray(cdf, dtype=np.int32)
        self.length = np.ascontiguousarray(length, dtype=np.int32)
        self.offset = np.ascontiguousarray(offset, dtype=np.int32)

    def tensors(self, device):
        return (torch.from_numpy(self.cdf.copy()).to(device), torch.from_numpy(self.length.copy()).to(device),
                torch.from_numpy(self.offset.copy()).to(device))


def _pmf_to_cdf(pmf, tail, pmf_length, max_length):
    """compressai `EntropyModel._pmf_to_cdf`: row i = quantised CDF of (pmf[i][:len_i], tail[i]), zero padded."""
    lib = L.load()
    rows = len(pmf_length)
    cdf = np.zeros((rows, max_length + 2), dtype=np.int32)
    for i in range(rows):
        n = int(pmf_length[i])
        prob = np.ascontiguousarray(np.concatenate([pmf[i, :n], tail[i:i + 1]]), dtype=np.float32)
        row = np.zeros(n + 2, dtype=np.int32)
        L.check(lib.tdvc_pmf_to_quantized_cdf(_np_ptr(prob), n + 1, PRECISION, _np_ptr(row)), "pmf_to_quantized_cdf")
        cdf[i, :n + 2] = row
    return cdf


def eb_tables(eb_packed):
    """compressai `EntropyBottleneck.update`.  eb_packed = (mats [C][33], biases [C][13], factors [C][12], medians,
    quantiles [C][3], target) of model._Packed (matrices already softplus'ed, factors tanh'ed, on the host at pack time)."""
    mats, biases, factors, _, quantiles, _ = (t.detach().cpu() for t in eb_packed)
    Cc = quantiles.shape[0]
    medians = quantiles[:, 1]
    minima = torch.clamp(torch.ceil(medians - quantiles[:, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(quantiles[:, 2] - medians).int(), min=0)
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]   # (C, 1, L)
    widths = (1, 3, 3, 3, 3, 1)
    mo = bo = fo = 0
    layers = []
    for i in range(5):
        fi, fn = widths[i], widths[i + 1]
        m = mats[:, mo:mo + fn * fi].reshape(Cc, fn, fi)
        bb = biases[:, bo:bo + fn].reshape(Cc, fn, 1)
        f = factors[:, fo:fo + fn].reshape(Cc, fn, 1) if i < 4 else None
        layers.append((m, bb, f))
        mo, bo, fo = mo + fn * fi, bo + fn, fo + (fn if i < 4 else 0)

    def logits(v):   # compressai `_logits_cumulative`
        for m, bb, f in layers:
            v = torch.matmul(m, v) + bb
            if f is not None:
                v = v + f * torch.tanh(v)
        return v

    lower, upper = logits(samples - 0.5), logits(samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = (torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:]))[:, 0]
    cdf = _pmf_to_cdf(pmf.numpy(), tail.numpy(), pmf_length.numpy(), max_length)
    return Tables(cdf, (pmf_length + 2).numpy(), (-minima).numpy())


_GC_TABLES = {}


def gc_tables(table=None):
    """compressai `GaussianConditional.update_scale_table(get_scale_table())` + `update()`: depends on nothing but the scale
    table and the tail mass, built once per process."""
    table = scale_table() if table is None else torch.as_tensor(table, dtype=torch.float32)
    key = tuple(table.tolist())
    if key in _GC_TABLES:
        return _GC_TABLES[key]
    multiplier = -statistics.NormalDist().inv_cdf(TAIL_MASS / 2)   # scipy.stats.norm.ppf in compressai
    center = torch.ceil(table * multiplier).int()
    pmf_length = 2 * center + 1
    max_length = int(pmf_length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    scale = table.unsqueeze(1).float()

    def phi(t):   # compressai `_standardized_cumulative`
        return 0.5 * torch.erfc(-(2 ** -0.5) * t)

    upper, lower = phi((0.5 - samples) / scale), phi((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = (2 * lower[:, :1])[:, 0]
    cdf = _pmf_to_cdf(pmf.numpy(), tail.numpy(), pmf_length.numpy(), max_length)
    _GC_TABLES[key] = Tables(cdf, (pmf_length + 2).numpy(), (-center).numpy())
    return _GC_TABLES[key]


def rans_encode(symbols, indexes, tables):
    """int32 numpy symbols / table indexes (same length, coding order) -> bytes."""
    lib = L.load()
    symbols = np.ascontiguousarray(symbols, dtype=np.int32).reshape(-1)
    indexes = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    if symbols.shape != indexes.shape:
        raise RuntimeError("rans_encode: symbols and indexes differ in length")
    cap = 8 * symbols.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = lib.tdvc_rans_encode_with_indexes(_np_ptr(symbols), _np_ptr(indexes), symbols.size, _np_ptr(tables.cdf),
                                          tables.cdf.shape[1], _np_ptr(tables.length), _np_ptr(tables.offset),
                                          tables.cdf.shape[0], _np_ptr(out), cap)
    if n < 0:
        L.check(int(n), "rans_encode_with_indexes")
    return out[:n].tobytes()


def rans_decode(data, indexes, tables):
    """Inverse of rans_encode given the same indexes -> int32 numpy symbols."""
    lib = L.load()
    indexes = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    buf = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(indexes.size, dtype=np.int32)
    L.check(lib.tdvc_rans_decode_with_indexes(_np_ptr(buf), buf.size, _np_ptr(indexes), indexes.size, _np_ptr(tables.cdf),
                                              tables.cdf.shape[1], _np_ptr(tables.length), _np_ptr(tables.offset),
                                              tables.cdf.shape[0], _np_ptr(out)), "rans_decode_with_indexes")
    return out


class CoderTables:
    """Tables of one coder (`update(force=True)`), built once per set of packed weights."""

    def __init__(self, W, cn, device):
        self.eb = eb_tables(W[f"{cn}.eb"])
        self.gc = gc_tables()
        self.scale_table = scale_table().to(device)


def _pinned(plan, name, like):
    key = ("pinned", name, tuple(like.shape), like.dtype)
    t = plan.bufs.get(key)
    if t is None:
        t = plan.bufs[key] = torch.empty(like.shape, dtype=like.dtype).pin_memory()
    return t


def launch_coding(plan, W, cn, tables, cluster=0, stream=None):
    """Device half of compressai `compress()` for coder `cn` ("mv" | "rs") on the latents the forward pass of `plan` just
    produced: symbols + table indexes of z and y, copied to pinned host memory on `stream`.  Returns a handle for
    finish_coding (the host half: rANS)."""
    lib = L.load()
    dev, N = plan.dev, plan.N
    stream = stream or torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    hy, wy, hz, wz = plan.H // 16, plan.W // 16, plan.H // 64, plan.W // 64
    y, z = plan.buf(f"{cn}.y", N, hy, wy, 128), plan.buf(f"{cn}.z", N, hz, wz, 128)
    params = plan.buf(f"{cn}.params", N, hy, wy, 256)
    med = W[f"{cn}.eb"][3]
    # ---- z: factorised prior, NCHW symbol order
    z_sym = plan.raw(f"{cn}.ac.z_sym", (N, 128, hz, wz), torch.int32)
    z_idx = plan.raw(f"{cn}.ac.z_idx", (N, 128, hz, wz), torch.int32)
    L.check(lib.tdvc_eb_symbols(z.ptr, z.ld, med.data_ptr(), N, hz * wz, 128, z_sym.data_ptr(), z_idx.data_ptr(), st),
            "eb_symbols")
    # ---- y: autoregressive pass, (h, w, c) symbol order
    ctx, e0, e2, e4 = W[f"{cn}.ctx"], W[f"{cn}.ep0"], W[f"{cn}.ep2"], W[f"{cn}.ep4"]
    y_hat = plan.raw(f"{cn}.ac.y_hat", (N, hy, wy, 128))
    y_sym = plan.raw(f"{cn}.ac.y_sym", (N, hy, wy, 128), torch.int32)
    y_idx = plan.raw(f"{cn}.ac.y_idx", (N, hy, wy, 128), torch.int32)
    if ctx.cin_pad != 128 or ctx.cout_pad != 256 or e0.cin_pad != 512 or e2.cin_pad < e0.cout_pad - 4 or e4.cout_pad != 256:
        raise RuntimeError("code_latents: unexpected entropy-parameter layer shapes")
    need = lib.tdvc_ar_code_workspace_bytes(N, e0.cout_pad, e2.cout_pad)
    ws = plan.raw(f"{cn}.ac.ws", ((need + 3) // 4,))
    p = L.ArParams(y=y.ptr, y_ld=y.ld, params=params.ptr, params_ld=params.ld, w_ctx=ctx.w.data_ptr(), b_ctx=ctx.b.data_ptr(),
                   w1=e0.w.data_ptr(), b1=e0.b.data_ptr(), c1=e0.cout, c1_pad=e0.cout_pad,
                   w2=e2.w.data_ptr(), b2=e2.b.data_ptr(), c2=e2.cout, c2_pad=e2.cout_pad,
                   w3=e4.w.data_ptr(), b3=e4.b.data_ptr(), scale_table=tables.scale_table.data_ptr(), n_scales=SCALES_LEVELS,
                   y_hat=y_hat.data_ptr(), symbols=y_sym.data_ptr(), indexes=y_idx.data_ptr(), N=N, H=hy, W=wy, C=128,
                   cluster=cluster)
    L.check(lib.tdvc_ar_code(C.byref(p), ws.data_ptr(), need, st), "ar_code")
    plan.launches += 2
    host = []
    with torch.cuda.stream(stream):
        for nme, t in (("y_sym", y_sym), ("y_idx", y_idx), ("z_sym", z_sym), ("z_idx", z_idx)):
            h = _pinned(plan, f"{cn}.ac.{nme}", t)
            h.copy_(t, non_blocking=True)
            host.append(h)
        done = torch.cuda.Event()
        done.record(stream)
    return {"host": host, "done": done, "tables": tables, "N": N, "shape": (hz, wz),
            "dev": {"y_symbols": y_sym, "y_indexes": y_idx, "y_hat": y_hat, "z_symbols": z_sym}}


def finish_coding(h, keep=False):
    """Host half: rANS over the symbols, one string per image (compressai codes batch items separately).
    Returns {"strings": [y_strings, z_strings], "shape": (h, w)} (+ clones of the device tensors if keep)."""
    h["done"].synchronize()
    hs = [t.numpy() for t in h["host"]]
    tables = h["tables"]
    y_strings = [rans_encode(hs[0][i], hs[1][i], tables.gc) for i in range(h["N"])]
    z_strings = [rans_encode(hs[2][i], hs[3][i], tables.eb) for i in range(h["N"])]
    out = {"strings": [y_strings, z_strings], "shape": h["shape"]}
    if keep:
        out.update({k: v.clone() for k, v in h["dev"].items()})
    return out


def code_latents(plan, W, cn, tables, cluster=0, keep=False):
    return finish_coding(launch_coding(plan, W, cn, tables, cluster=cluster), keep=keep)


# ----------------------------------------------------------------------------------------------- bitstream container
_MAGIC = b"TDVC"


def pack_frame(coded):
    """`net.last_coded` of one P-frame -> bytes.  The reference only sketches a container (dead code, reference
    tools/utils/encoder.py:61-68: per string a big-endian shape header, a length and the payload); this one follows the
    sketch with 32-bit lengths (a 1920x1024 motion string is ~1 MB, beyond the sketch's uint16):
    magic "TDVC" | >B n_coders | per coder: >4s name | >2I hyper-latent shape (h, w) | >B n_lists | per list: >I n_strings |
    per string: >I length | payload."""
    import struct
    out = [_MAGIC, struct.pack(">B", len(coded))]
    for name, enc in coded.items():
        out.append(struct.pack(">4s2IB", name.encode()[:4].ljust(4), int(enc["shape"][0]), int(enc["shape"][1]),
                               len(enc["strings"])))
        for lst in enc["strings"]:
            out.append(struct.pack(">I", len(lst)))
            for s_ in lst:
                out.append(struct.pack(">I", len(s_)))
                out.append(bytes(s_))
    return b"".join(out)


def unpack_frame(data):
    """Inverse of pack_frame -> {name: {"strings": [[bytes, ...], ...], "shape": (h, w)}}."""
    import struct
    if data[:4] != _MAGIC:
        raise RuntimeError("unpack_frame: not a tdvc_b200 frame container")
    pos = 4
    (n,) = struct.unpack_from(">B", data, pos)
    pos += 1
    out = {}
    for _ in range(n):
        name, h, w, nl = struct.unpack_from(">4s2IB", data, pos)
        pos += struct.calcsize(">4s2IB")
        lists = []
        for _ in range(nl):
            (ns,) = struct.unpack_from(">I", data, pos)
            pos += 4
            strs = []
            for _ in range(ns):
                (ln,) = struct.unpack_from(">I", data, pos)
                pos += 4
                if pos + ln > len(data):
                    raise RuntimeError("unpack_frame: truncated container")
                strs.append(bytes(data[pos:pos + ln]))
                pos += ln
            lists.append(strs)
        out[name.decode().strip()] = {"strings": lists, "shape": (h, w)}
    if pos != len(data):
        raise RuntimeError("unpack_frame: trailing bytes")
    return out
