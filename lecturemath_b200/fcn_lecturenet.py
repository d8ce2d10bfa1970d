"""FCN_LectureNet drop-in (R/AccessMath/lecturenet_v1/FCN_lecturenet.py) whose inference runs on hand-written
sm_100a kernels (csrc/fcn_conv.cu: TMA + tcgen05 implicit GEMM; csrc/fcn_misc.cu: glue) through the C ABI.

Kept from the reference's surface: FCN_LectureNet.CreateFromConfig(config, in_channels, reconstruction_mode)
(:620-659), state_dict()/load_state_dict() with the reference's parameter names (weight interchange format),
eval(), cuda(), forward(x) -> (logit, text_logit, rec) (:364-403), binarize(PIL_image, return_others, force_binary,
binary_treshold, apply_sigmoid) (:430-505), prepare_image (:607-618).

The torch.nn modules below are parameter CONTAINERS only (same construction order as the reference, so the same
seed yields the same random-init weights); no torch op runs in forward/binarize."""
import ctypes
import json
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------------
# parameter container with the reference's names
def _block(cin, cout, k, act):
    layers = [nn.Conv2d(cin, cout, stride=1, kernel_size=k, padding=(k - 1) // 2), nn.BatchNorm2d(cout)]
    if act == "gelu":
        layers.append(nn.GELU())
    elif act == "tanh":
        layers.append(nn.Tanh())
    seq = nn.Sequential(*layers)
    nn.init.xavier_normal_(seq[0].weight)
    return seq


class _Params(nn.Module):
    """Builds parameters in the order FCN_LectureNet.__init__ does (:17-162) so seeds reproduce its init."""

    def __init__(self, ch, down, mid, ups, upc, k, pm1, pm2, pk):
        super().__init__()
        cin = ch
        for i, c in enumerate(down, 1):
            setattr(self, "conv_down_block_%d" % i, _block(cin, c, k, "gelu"))
            cin = c
        self.mid_block = _block(down[4], mid, k, "gelu")
        cin = mid
        for lvl in range(5, 0, -1):
            t = nn.ConvTranspose2d(cin, ups[lvl - 1], 2, padding=0, stride=2)
            setattr(self, "transposed_conv_%d" % lvl, t)
            setattr(self, "upsample_block_%d" % lvl, nn.Sequential(nn.BatchNorm2d(ups[lvl - 1]), nn.GELU()))
            nn.init.xavier_normal_(t.weight)
            setattr(self, "conv_up_block_%d" % lvl, _block(ups[lvl - 1] + down[lvl - 1], upc[lvl - 1], k, "gelu"))
            cin = upc[lvl - 1]
        # set_main_branches (:164-201): pixel branch then text-mask branch, kernel = pixel kernel size
        self.conv_pixels_1 = _block(ch + upc[0], pm1, pk, "gelu")
        self.conv_pixels_2 = _block(ch + pm1, pm2, pk, "gelu")
        self.conv_out = _block(ch + pm2, 1, pk, None)
        self.conv_text_mask_out = _block(upc[0], 1, pk, None)
        self.conv_reconstruct = _block(upc[0], 3, k, "tanh")


# ----------------------------------------------------------------------------------------------------
# ctypes mirrors of include/accessmath_b200.h
class ConvSeg(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("C", ctypes.c_int), ("Wp", ctypes.c_int), ("Hbuf", ctypes.c_int), ("x_off", ctypes.c_int),
                ("rowrun", ctypes.c_int), ("S", ctypes.c_int), ("run_len", ctypes.c_int), ("KW", ctypes.c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("nseg", ctypes.c_int), ("seg", ConvSeg * 2), ("weights", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("KH", ctypes.c_int), ("padY", ctypes.c_int), ("RT", ctypes.c_int), ("YT", ctypes.c_int),
                ("nR", ctypes.c_int), ("Hin", ctypes.c_int), ("batch", ctypes.c_int),
                ("NT", ctypes.c_int), ("Ntot", ctypes.c_int), ("Ntot_pad", ctypes.c_int),
                ("out", ctypes.c_void_p), ("out_f32", ctypes.c_int), ("out_H", ctypes.c_int), ("out_W", ctypes.c_int),
                ("out_sn", ctypes.c_longlong), ("out_sy", ctypes.c_longlong),
                ("out_sx", ctypes.c_int), ("out_padx", ctypes.c_int), ("out_coff", ctypes.c_int),
                ("Cout", ctypes.c_int), ("Sy", ctypes.c_int), ("Sx", ctypes.c_int), ("act", ctypes.c_int), ("flags", ctypes.c_int), ("in_ystep", ctypes.c_int),
                ("pool_out", ctypes.c_void_p), ("pool_H", ctypes.c_int), ("pool_W", ctypes.c_int),
                ("pool_sn", ctypes.c_longlong), ("pool_sy", ctypes.c_longlong), ("pool_sx", ctypes.c_int), ("pool_padx", ctypes.c_int),
                ("epi_mode", ctypes.c_int), ("frames", ctypes.c_void_p), ("diff_out", ctypes.c_void_p), ("diff_C", ctypes.c_int),
                ("diff_pad", ctypes.c_int), ("text_out", ctypes.c_void_p), ("rec_out", ctypes.c_void_p), ("bits_out", ctypes.c_void_p),
                ("bits_wpr", ctypes.c_int), ("threshold", ctypes.c_int)]


AM_EPI_PLAIN, AM_EPI_HEADS, AM_EPI_THRESHOLD = 0, 1, 2
AM_BIND_FRAMES, AM_BIND_TEXT_OUT, AM_BIND_REC_OUT, AM_BIND_OUT, AM_BIND_THRESHOLD = 0, 1, 2, 3, 4


def fold_bn(w, b, bn_w, bn_b, mean, var, out_dim=0):
    """Conv + eval BatchNorm -> conv:  w' = w*g/sqrt(v+eps), b' = (b-mu)*g/sqrt(v+eps)+beta  (SURVEY appendix A)."""
    scale = bn_w.double() / torch.sqrt(var.double() + BN_EPS)
    shape = [1] * w.dim()
    shape[out_dim] = -1
    return (w.double() * scale.view(shape)).float(), ((b.double() - mean.double()) * scale + bn_b.double()).float()


def choose_nt(ntot):
    if ntot <= 256:
        return ((ntot + 15) // 16) * 16
    best = None
    for nt in (256, 192, 128):
        pad = ((ntot + nt - 1) // nt) * nt
        if best is None or pad < best[1]:
            best = (nt, pad)
    return best[0]


def choose_tile(nr, h, kh):
    best = None
    for rt, yt in ((8, 16), (16, 8), (32, 4), (64, 2), (128, 1)):
        area = math.ceil(nr / rt) * rt * math.ceil(h / yt) * yt
        score = area * (1.0 + 0.25 * (kh - 1) / yt)
        if best is None or score < best[0]:
            best = (score, rt, yt)
    return best[1], best[2]


def pack_weights(w, bias, segs, S, rowrun, NT, Sy=1):
    """Re-pack a folded filter for the row-run implicit GEMM (csrc/fcn_conv.cu header comment).

    w    : fp32 [Nrows][Cin][KH][KW]  (Nrows = Cout, or 4*Cout for the transposed conv's (sy,sx,co))
    segs : list of (C_buf, index tensor mapping buffer channel -> w input channel, -1 = zero padding)
    S, Sy: output pixels packed per GEMM row in x and (2-D packing) in y; GEMM column n = ((sy*S)+sx)*Nrows + co and the
           filter becomes a zero-padded Toeplitz block over KW+S-1 horizontal and KH+Sy-1 vertical taps
    ->   bf16 [chunks*KHe*Ntot_pad][64], fp32 bias [Ntot_pad], Ntot, Ntot_pad        (KHe = KH + Sy - 1)
    """
    nrows, _, KH, KW = w.shape
    KHe = KH + Sy - 1
    ntot = Sy * S * nrows
    ntot_pad = ((ntot + NT - 1) // NT) * NT
    blocks = []
    for C, cmap in segs:
        cmap = torch.as_tensor(cmap, dtype=torch.long)
        wseg = torch.zeros((nrows, C, KH, KW), dtype=torch.float32)
        valid = cmap >= 0
        wseg[:, valid] = w[:, cmap[valid]]
        if rowrun:
            J = KW + S - 1
            run = torch.zeros((Sy, S, nrows, KHe, J, C), dtype=torch.float32)
            for sy in range(Sy):
                for sx in range(S):
                    run[sy, sx, :, sy:sy + KH, sx:sx + KW, :] = wseg.permute(0, 2, 3, 1)    # [co][dy][tap][c]
            K = J * C
            mat = run.permute(3, 0, 1, 2, 4, 5).reshape(KHe, ntot, K)            # [dy'][n=(sy,sx,co)][k=(j,c)]
            nck = (K + 63) // 64
            full = torch.zeros((KHe, ntot_pad, nck * 64), dtype=torch.float32)
            full[:, :ntot, :K] = mat
            blocks.append(full.view(KHe, ntot_pad, nck, 64).permute(2, 0, 1, 3))  # [ck][dy'][n][64]
        else:
            assert S == 1 and Sy == 1
            nck = (C + 63) // 64
            full = torch.zeros((KW, KH, ntot_pad, nck * 64), dtype=torch.float32)
            full[:, :, :ntot, :C] = wseg.permute(3, 2, 0, 1)                     # [kx][dy][co][c]
            blocks.append(full.view(KW, KH, ntot_pad, nck, 64).permute(0, 3, 1, 2, 4).reshape(KW * nck, KH, ntot_pad, 64))
    packed = torch.cat(blocks, 0).reshape(-1, 64).to(torch.bfloat16).contiguous()
    b = torch.zeros(ntot_pad, dtype=torch.float32)
    b[:ntot] = bias.repeat(Sy * S)
    return packed, b, ntot, ntot_pad


TUNED_PATH = os.environ.get("AM_B200_TUNED") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned_plans.json")  # env: tuning A/B only
_TUNED = None


def arch_signature(sd):
    """Key of the tuned table: everything the per-layer GEMM shapes depend on besides (batch, H, W) -- channel widths and kernel sizes
    (the FCN_BINARIZER_NET_* keys are configurable; a table measured for one architecture must not steer another)."""
    w = lambda n: sd[n].shape
    down = ".".join(str(w("conv_down_block_%d.0.weight" % i)[0]) for i in range(1, 6))
    ups = ".".join(str(w("transposed_conv_%d.weight" % i)[1]) for i in range(1, 6))
    upc = ".".join(str(w("conv_up_block_%d.0.weight" % i)[0]) for i in range(1, 6))
    return "c%d_d%s_m%d_t%s_u%s_k%d_p%d_pm%d.%d" % (w("conv_down_block_1.0.weight")[1], down, w("mid_block.0.weight")[0], ups, upc,
                                                    w("conv_down_block_1.0.weight")[-1], w("conv_pixels_1.0.weight")[-1],
                                                    w("conv_pixels_1.0.weight")[0], w("conv_pixels_2.0.weight")[0])


def tuned_table():
    """{architecture signature: {"B<batch>_<H>x<W>": {layer name: [S, Sy, NT, MT]}}} measured on a B200 by tools/autotune_fcn.py
    (per-layer CUDA-event timings of every feasible packing / tiling); layers or shapes without an entry fall back to the cycle model."""
    global _TUNED
    if _TUNED is None:
        _TUNED = {}
        if os.path.exists(TUNED_PATH):
            with open(TUNED_PATH) as f:
                _TUNED = json.load(f)
    return _TUNED


SMEM_BUDGET = 226 * 1024
AM_CONV_NO_MT2, AM_CONV_FORCE_MT2, AM_CONV_FORCE_MT4, AM_CONV_CTA_PAIR = 2, 4, 8, 16
MT_PAIR = 22            # configuration code: CTA pair (cta_group::2), two M-tiles per CTA, four per work item
MT_FLAGS = {1: AM_CONV_NO_MT2, 2: AM_CONV_FORCE_MT2, 4: AM_CONV_FORCE_MT4, MT_PAIR: AM_CONV_CTA_PAIR}
MT_OPTIONS = (1, 2, 4, MT_PAIR)
L2_BYTES_PER_CLK_SM = 42.0        # ~6300 B/clk chip-wide L2->SM throughput / 148 SMs (B300_MICROARCH.md)
EPI_CLK_PER_COL = 43.0            # epilogue cycles per accumulator column of a 128-row tile (16 warps; ~26 SASS instr/element)
EPI_CLK_PER_TILE = 600.0          # fixed epilogue cost per tile (barrier round trip, tile decode)
ISSUE_CLK_PER_MMA = 110.0         # an issuing warp is latency bound: ~100-120 clk per tcgen05.mma measured on the narrow layers (r01 profiles)
ISSUE_CLK_PER_DY = (30.0, 150.0)  # per (chunk, dy) step: resident / streamed weights (mbarrier try_wait + commit)
ISSUE_CLK_PER_CHUNK = 120.0       # per A chunk: full-barrier wait + commit


def nt_candidates(ntot):
    """UMMA N per CTA: one block when it fits, else / also equal splits (more N blocks = smaller accumulators, which is
    what lets two M-tiles AND two accumulator stages share the 512 TMEM columns)."""
    r16 = ((ntot + 15) // 16) * 16
    out = []
    for nt in (256, 192, 128, 96, 64, 48, 32, 16):
        if nt > r16:
            continue
        pad = ((ntot + nt - 1) // nt) * nt
        if nt == r16 or (pad - ntot) * 10 <= ntot and nt >= 64:
            out.append(nt)
    if r16 <= 256 and r16 not in out:
        out.append(r16)
    return out


def layer_cost(nr, h, ntot, runs, kh, batch, nt=None, mt=None, sy=1, n_sm=148):
    """Cycle model of k_conv_gemm (csrc/fcn_conv.cu) for one layer: GEMM width ntot, per-segment run lengths `runs`
    (elements of K per vertical tap), UMMA N `nt` and `mt` M-tiles per work item.  Mirrors conv_prepare()'s feasibility
    rules (resident weights, accumulator stages).  Calibrated against profiles/ (r01): narrow layers are bound by the
    MMA-issuing warps, single-stage accumulators serialise main loop and epilogue.  Returns None when infeasible."""
    nt = nt or choose_nt(ntot)
    pair = mt == MT_PAIR
    if pair:                                      # per CTA it is the mt = 2 kernel with half of every weight tile
        mt = 2
        if nt % 16:
            return None
    ntot_pad = ((ntot + nt - 1) // nt) * nt
    nnb = ntot_pad // nt
    if sy == 2:                                   # 2-D packing: `h` image rows -> ceil(h/2) GEMM row units, kh + 1 Toeplitz taps, RT = 8
        h, kh, (rt, yt) = (h + 1) // 2, kh + 1, (8, 16)
    else:
        rt, yt = choose_tile(nr, h, kh)
    chunks = sum((r + 63) // 64 for r in runs)
    bytes_a, bytes_b = (sy * (yt - 1) + kh) * rt * 128, nt * 128
    bytes_b_all = chunks * kh * bytes_b
    fixed = 2560
    n_mtiles = math.ceil(nr / rt) * math.ceil(h / yt) * batch
    ntc = 32
    while ntc < nt:
        ntc *= 2
    if mt is None:
        mt = 2 if (n_mtiles * nnb >= 4 * n_sm and 2 * ntc <= 512) else 1
    if mt * ntc > 512:
        return None
    if pair:
        bytes_b //= 2
        bytes_b_all //= 2
    resident = not pair and nnb == 1 and fixed + bytes_b_all + (2 * mt if mt >= 2 else 3) * bytes_a <= SMEM_BUDGET
    if not resident and fixed + 2 * mt * bytes_a + 2 * bytes_b > SMEM_BUDGET:
        return None
    acc_stages = 2 if 2 * mt * ntc <= 512 else 1
    t_mma = max(nt / 2.0, (4096 + (16 if pair else 32) * nt) / 128.0)  # tensor floor vs smem operand read, per K=16 step
    mma = 0.0
    for r in runs:
        nck = (r + 63) // 64
        for ck in range(nck):
            ks = 4 if ck < nck - 1 else ((r - ck * 64) + 15) // 16
            per_dy = max(mt * ks * t_mma, ks * ISSUE_CLK_PER_MMA + ISSUE_CLK_PER_DY[0 if resident else 1])   # one issuer per M-tile
            mma += kh * per_dy + ISSUE_CLK_PER_CHUNK
    l2 = (mt * chunks * bytes_a + (0 if resident else bytes_b_all)) / L2_BYTES_PER_CLK_SM
    epi = mt * (EPI_CLK_PER_COL * nt + EPI_CLK_PER_TILE) * (16.0 / 12.0 if mt == 4 else 1.0)      # MT = 4 leaves 12 epilogue warps
    item = max(mma, l2, epi) if acc_stages == 2 else max(mma + epi, l2)
    n_items = math.ceil(n_mtiles / mt) * nnb
    if pair:
        n_items = math.ceil(n_mtiles / 4) * nnb * 2                  # per-CTA items, strided over n_sm / 2 clusters
    total = math.ceil(n_items / n_sm) * item
    return {"clk": total, "item": item, "mma": mma, "l2": l2, "epi": epi, "resident": resident, "mt": mt, "nt": nt,
            "acc_stages": acc_stages, "items": n_items}


class _Buf:
    """NHWC bf16 activation buffer with physical zero padding in x."""

    def __init__(self, B, H, W, C, pad, device, dtype=torch.bfloat16):
        self.B, self.H, self.W, self.C, self.pad = B, H, W, C, pad
        self.Wp = W + 2 * pad
        n = B * H * self.Wp * C
        self.t = torch.zeros(n + 256, dtype=dtype, device=device)      # slack: TMA boxes never leave the allocation
        self.ptr = self.t.data_ptr()

    def view(self):
        return self.t[:self.B * self.H * self.Wp * self.C].view(self.B, self.H, self.Wp, self.C)[:, :, self.pad:self.pad + self.W]


class FCNPlan:
    """All device buffers, packed weights and kernel descriptors for one (batch, H, W)."""

    def __init__(self, net, B, H, W, device, rowrun=True, overrides=None):
        self.ov = dict(overrides or {})      # planner overrides for tests / tuning, e.g. {"sy": 2} forces 2-D packing where legal
        self.lib = _lib.load()          # building a plan needs no device; run() does (and fails loudly without one)
        self.B, self.H, self.W, self.device, self.rowrun = B, H, W, device, rowrun
        sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        self.arch = arch_signature(sd)
        k = sd["conv_down_block_1.0.weight"].shape[-1]
        pk = sd["conv_pixels_1.0.weight"].shape[-1]
        if k > pk:
            raise ValueError("FCN_BINARIZER_NET_KERNEL_SIZE (%d) larger than FCN_BINARIZER_NET_PIXEL_KERNEL_SIZE (%d): the text-mask and "
                             "reconstruction heads share one %dx%d GEMM, which cannot hold the larger filter" % (k, pk, pk, pk))
        p3, p7 = (k - 1) // 2, (pk - 1) // 2
        down = [sd["conv_down_block_%d.0.weight" % i].shape[0] for i in range(1, 6)]
        mid = sd["mid_block.0.weight"].shape[0]
        ups = [sd["transposed_conv_%d.weight" % i].shape[1] for i in range(1, 6)]
        upc = [sd["conv_up_block_%d.0.weight" % i].shape[0] for i in range(1, 6)]
        pm1, pm2 = sd["conv_pixels_1.0.weight"].shape[0], sd["conv_pixels_2.0.weight"].shape[0]
        for c in down + [mid] + ups + upc + [pm1, pm2]:
            if c % 8:
                raise ValueError("channel widths must be multiples of 8 for the NHWC bf16 kernels (got %d)" % c)
        hs, ws = [H], [W]
        for _ in range(5):
            hs.append(hs[-1] // 2); ws.append(ws[-1] // 2)
        if hs[5] < 1 or ws[5] < 1:
            raise ValueError("image too small for five pooling levels")
        mk = lambda lvl, C, pad: _Buf(B, hs[lvl], ws[lvl], C, pad, device)
        self.frames = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=device)
        self.x0 = mk(0, 8, p3)
        self.d = [mk(i, down[i], p3) for i in range(5)]                  # conv_down outputs (skip connections)
        self.p = [mk(i + 1, down[i], p3) for i in range(5)]              # pooled
        self.mid = mk(5, mid, 0)
        self.t = [mk(i, ups[i], p3) for i in range(5)]                   # transposed-conv outputs (level i)
        self.u = [mk(i, upc[i], 0 if i > 0 else p7) for i in range(5)]   # conv_up outputs; u[0] = x_up1
        # `diff` has 3 channels: 4 per pixel when every consumer reads it through a row-run view whose row stride S * 4 * 2 bytes is a
        # multiple of 16 (TMA), i.e. S even -- the 7x7 layers then pad 35 -> 36 input channels instead of 35 -> 40; else 8
        dc = 4
        if not rowrun or self.ov.get("diff_c") == 8:
            dc = 8
        else:
            for name, cout, cin, cap in (("conv_pixels_1", pm1, upc[0], None), ("conv_pixels_2", pm2, pm1, None), ("conv_out", 1, pm2, 64)):
                if self._pick_config(name, W, H, cout, [cin, 4], pk, pk, cap)[0] % 2:
                    dc = 8
        self.diff = mk(0, dc, p7)
        self.px1, self.px2 = mk(0, pm1, p7), mk(0, pm2, p7)
        self.logits = torch.zeros((B, H, W), dtype=torch.float32, device=device)
        self.text_logit = torch.zeros((B, H, W), dtype=torch.float32, device=device)
        self.rec = torch.zeros((B, H, W, 3), dtype=torch.float32, device=device)
        self.wpr = self.lib.am_words_per_row(W)
        self.bits = torch.zeros((B, H, self.wpr), dtype=torch.int32, device=device)
        self._borders_filled = False
        self._cplans = {}       # op index -> am_conv_plan handle
        self.keep = []          # keeps packed weights alive
        self.ops = []           # (kind, payload)
        self.op_flops = {}      # op index -> algorithmic FLOPs per frame of that conv launch
        self.specs = {}         # layer name -> what _conv / _tconv were given (+ op index, chosen cfg): autotuning builds variants
        self.flops = 0

        def cbn(name):
            return fold_bn(sd[name + ".0.weight"], sd[name + ".0.bias"], sd[name + ".1.weight"], sd[name + ".1.bias"],
                           sd[name + ".1.running_mean"], sd[name + ".1.running_var"])

        ident = lambda c: list(range(c))
        # encoder
        src = self.x0
        for i in range(5):
            w, b = cbn("conv_down_block_%d" % (i + 1))
            cmap = [0, 1, 2] + [-1] * 5 if i == 0 else ident(src.C)
            if not self._conv("conv_down_block_%d" % (i + 1), w, b, [(src, cmap)], self.d[i], act=1, pool_dst=self.p[i]):
                self.ops.append(("pool", (self.d[i], self.p[i])))      # not fusable in this configuration (S / Sy packing, RT = 32)
            src = self.p[i]
        w, b = cbn("mid_block")
        self._conv("mid_block", w, b, [(src, ident(src.C))], self.mid, act=1)
        # decoder
        src = self.mid
        for lvl in range(4, -1, -1):
            name = "transposed_conv_%d" % (lvl + 1)
            bn = "upsample_block_%d.0" % (lvl + 1)
            wt, bt = fold_bn(sd[name + ".weight"], sd[name + ".bias"], sd[bn + ".weight"], sd[bn + ".bias"],
                             sd[bn + ".running_mean"], sd[bn + ".running_var"], out_dim=1)
            self._tconv(name, wt, bt, src, self.t[lvl])
            w, b = cbn("conv_up_block_%d" % (lvl + 1))
            cu = self.t[lvl].C
            self._conv("conv_up_block_%d" % (lvl + 1), w, b, [(self.t[lvl], ident(cu)), (self.d[lvl], [cu + c for c in range(self.d[lvl].C)])],
                       self.u[lvl], act=1)
            src = self.u[lvl]
        # heads: text mask (pk x pk, 1 ch) and reconstruction (k x k, 3 ch) share one pk x pk GEMM with 4 columns
        wt_, bt_ = cbn("conv_text_mask_out")
        wr_, br_ = cbn("conv_reconstruct")
        wh = torch.zeros((4, upc[0], pk, pk))
        wh[0] = wt_[0]
        o = (pk - k) // 2
        wh[1:4, :, o:o + k, o:o + k] = wr_
        f0 = self.flops
        # ... whose epilogue applies tanh / sigmoid and writes diff = (x0 - rec) * sigmoid(text) (:370-377) directly (AM_EPI_HEADS)
        self._conv("heads", wh, torch.cat([bt_, br_]), [(self.u[0], ident(upc[0]))], None, act=0, cap=16, epi=AM_EPI_HEADS)
        self.flops = f0 + 2 * H * W * upc[0] * (pk * pk + 3 * k * k)     # algorithmic: 7x7x1 + 3x3x3, not the padded GEMM
        self.op_flops[len(self.ops) - 1] = self.flops - f0
        self.heads_op = len(self.ops) - 1
        dmap = [0, 1, 2] + [-1] * (dc - 3)
        w, b = cbn("conv_pixels_1")       # reference input order: (diff 0..2, x_up1)  (:383)
        self._conv("conv_pixels_1", w, b, [(self.u[0], [3 + c for c in range(upc[0])]), (self.diff, dmap)], self.px1, act=1)
        w, b = cbn("conv_pixels_2")
        self._conv("conv_pixels_2", w, b, [(self.px1, [3 + c for c in range(pm1)]), (self.diff, dmap)], self.px2, act=1)
        w, b = cbn("conv_out")
        # conv_out: sigmoid-threshold + bit-packing in the epilogue (AM_EPI_THRESHOLD) when a GEMM row holds whole 16-pixel runs
        self._conv("conv_out", w, b, [(self.px2, [3 + c for c in range(pm2)]), (self.diff, dmap)], None, act=0, cap=64, f32_out=self.logits,
                   epi=AM_EPI_THRESHOLD)
        self.out_op = len(self.ops) - 1
        self.fused_threshold = self.ops[self.out_op][1].epi_mode == AM_EPI_THRESHOLD
        if not self.fused_threshold:
            self.ops.append(("threshold", None))

    @staticmethod
    def count_flops(net, H, W):
        """Algorithmic FLOPs per frame: 2 x MACs of the 16 conv + 5 transposed-conv layers, no padding."""
        sd = net.state_dict()
        hs, ws = [H], [W]
        for _ in range(5):
            hs.append(hs[-1] // 2); ws.append(ws[-1] // 2)
        macs = 0
        for i in range(1, 6):
            macs += hs[i - 1] * ws[i - 1] * sd["conv_down_block_%d.0.weight" % i][0].numel() * sd["conv_down_block_%d.0.weight" % i].shape[0]
            macs += hs[i - 1] * ws[i - 1] * sd["conv_up_block_%d.0.weight" % i].numel()
            macs += hs[i] * ws[i] * sd["transposed_conv_%d.weight" % i].numel()
        macs += hs[5] * ws[5] * sd["mid_block.0.weight"].numel()
        for n in ("conv_text_mask_out", "conv_reconstruct", "conv_pixels_1", "conv_pixels_2", "conv_out"):
            macs += H * W * sd[n + ".0.weight"].numel()
        return 2 * macs

    # -------------------------------------------------------------------------------------------------
    def _candidates(self, width, height, cout, seg_cs, KW, KH, cap=None):
        """Every feasible (S, Sy, NT, MT) with its modelled cycles, best first: S / Sy = output pixels per GEMM row in x / y,
        NT = UMMA N per CTA, MT = M-tiles per work item."""
        out = []
        sy_options = (1, 2) if KH > 1 and height >= 2 else (1,)
        if "sy" in self.ov and self.ov["sy"] in sy_options:
            sy_options = (self.ov["sy"],)
        for sy in sy_options:
            s = 1
            while s <= 32 and (cap is None or s * sy <= cap):
                if s == 1 or (width % s == 0 and s * sy * cout <= 256):
                    runs = [(KW + s - 1) * c_ for c_ in seg_cs]
                    for nt in nt_candidates(s * sy * cout):
                        for mt in MT_OPTIONS:
                            c = layer_cost(width // s, height, s * sy * cout, runs, KH, self.B, nt=nt, mt=mt, sy=sy)
                            if c is not None:
                                out.append((c["clk"], s, sy, nt, mt))
                s *= 2
        return out

    def _pick_config(self, name, width, height, cout, seg_cs, KW, KH, cap=None):
        """Configuration of layer `name`: explicit override, else the measured table (tuned_plans.json), else the model."""
        if not self.rowrun:
            return 1, 1, choose_nt(cout), None
        forced = self.ov.get("cfg", {}).get(name)
        if forced is not None:
            return tuple(forced)
        cands = self._candidates(width, height, cout, seg_cs, KW, KH, cap)
        if "sy" not in self.ov and "mt" not in self.ov and not self.ov.get("no_tuned"):
            tuned = tuned_table().get(self.arch, {}).get("B%d_%dx%d" % (self.B, self.H, self.W), {}).get(name)
            if tuned is not None and any(tuple(c[1:]) == tuple(tuned) for c in cands):      # only a configuration that is feasible HERE
                return tuple(tuned)
        best = None
        if "mt" in self.ov and any(c[4] == self.ov["mt"] for c in cands):       # tests: force the M-tiles per work item where feasible
            cands = [c for c in cands if c[4] == self.ov["mt"]]
        for c in cands:
            if best is None or c[0] < best[0] * 0.97:            # near ties: keep the earlier (less padding)
                best = c
        return best[1], best[2], best[3], best[4]

    def _conv(self, name, w, b, srcs, dst, act, cap=None, f32_out=None, pool_dst=None, epi=AM_EPI_PLAIN):
        """pool_dst: buffer for MaxPool2d(2) of the output; returns True when the pool was fused into this conv's epilogue.
        epi: fused epilogue of the fp32 layers (AM_EPI_HEADS / AM_EPI_THRESHOLD, include/accessmath_b200.h)."""
        nrows, cin_total, KH, KW = w.shape
        first = srcs[0][0]
        cfg = self._pick_config(name, first.W, first.H, nrows, [buf.C for buf, _ in srcs], KW, KH, cap)
        self.specs[name] = dict(kind="conv", w=w, b=b, srcs=srcs, dst=dst, act=act, cap=cap, f32_out=f32_out, op=len(self.ops), cfg=cfg,
                                pool_dst=pool_dst, epi=epi)
        d, keep = self._conv_desc(w, b, srcs, dst, act, f32_out, cfg, pool_dst, epi)
        self.keep += keep
        self.ops.append(("conv", d))
        fused = bool(d.pool_out)
        self.op_flops[len(self.ops) - 1] = 2 * first.H * first.W * nrows * cin_total * KH * KW
        self.flops += 2 * first.H * first.W * nrows * cin_total * KH * KW
        return fused

    def _conv_desc(self, w, b, srcs, dst, act, f32_out, cfg, pool_dst=None, epi=AM_EPI_PLAIN):
        """Kernel descriptor (+ the packed tensors it points to) of one convolution for configuration cfg = (S, Sy, NT, MT)."""
        nrows, cin_total, KH, KW = w.shape
        first = srcs[0][0]
        Hin, Win = first.H, first.W
        S, Sy, NT, MT = cfg
        packed, bias, ntot, ntot_pad = pack_weights(w, b, [(buf.C, cmap) for buf, cmap in srcs], S, self.rowrun, NT, Sy)
        Himg, KHc = Hin, KH
        Hin, KH = (Himg + Sy - 1) // Sy, KHc + Sy - 1          # GEMM row units per frame column, Toeplitz-extended vertical taps
        packed, bias = packed.to(self.device), bias.to(self.device)
        d = ConvDesc()
        d.nseg = len(srcs)
        for i, (buf, _) in enumerate(srcs):
            g = d.seg[i]
            g.ptr, g.C, g.Wp, g.Hbuf = buf.ptr, buf.C, buf.Wp, buf.H
            g.x_off = buf.pad - (KW - 1) // 2
            assert g.x_off >= 0
            g.rowrun, g.S, g.run_len, g.KW = int(self.rowrun), S, (KW + S - 1) * buf.C, KW
        d.weights, d.bias = packed.data_ptr(), bias.data_ptr()
        d.KH, d.padY, d.in_ystep = KH, (KHc - 1) // 2, Sy
        d.nR, d.Hin, d.batch = Win // S, Hin, self.B
        d.RT, d.YT = (8, 16) if Sy == 2 else choose_tile(d.nR, Hin, KH)
        d.NT, d.Ntot, d.Ntot_pad = NT, ntot, ntot_pad
        if epi == AM_EPI_THRESHOLD and (nrows != 1 or S % 16 or Win % 16 or not self.rowrun):
            epi = AM_EPI_PLAIN                                            # legacy path: fp32 logits + k_threshold_pack
        d.epi_mode = epi
        if epi != AM_EPI_PLAIN:                                           # pixel geometry: element = pixel of the fp32 [B][H][W] planes
            d.out, d.out_f32 = (f32_out.data_ptr() if f32_out is not None else None), 1
            d.out_H, d.out_W = Himg, Win
            d.out_sx, d.out_sy, d.out_sn = 1, Win, Himg * Win
            d.out_padx, d.out_coff = 0, 0
            if epi == AM_EPI_HEADS:
                d.frames = self.frames.data_ptr()
                d.diff_out, d.diff_C, d.diff_pad = self.diff.ptr, self.diff.C, self.diff.pad
            else:
                d.bits_out, d.bits_wpr, d.threshold = self.bits.data_ptr(), self.wpr, 128
        elif f32_out is not None:
            d.out, d.out_f32 = f32_out.data_ptr(), 1
            d.out_H, d.out_W = Himg, Win
            d.out_sx = nrows
            d.out_sy = Win * nrows
            d.out_sn = Himg * Win * nrows
            d.out_padx, d.out_coff = 0, 0
        else:
            d.out, d.out_f32 = dst.ptr, 0
            d.out_H, d.out_W = dst.H, dst.W
            d.out_sx, d.out_sy, d.out_sn = dst.C, dst.Wp * dst.C, dst.H * dst.Wp * dst.C
            d.out_padx, d.out_coff = dst.pad, 0
        d.Cout, d.Sy, d.Sx, d.act = nrows, Sy, S, act
        d.flags = 0 if MT is None else MT_FLAGS[MT]
        # MaxPool2d(2) fused into the epilogue where the 2x2 block of a pixel sits in four lanes of one warp (csrc/fcn_conv.cu epi_unit)
        fuse = False
        if pool_dst is not None and f32_out is None and nrows % 16 == 0 and self.rowrun and not self.ov.get("no_fused_pool"):
            if S == 1 and Sy == 1 and d.RT <= 16:
                # (level 2, K = 432, is epilogue bound: the fused pool costs its conv 0.11 ms and saves a 0.15 ms pass; level 1 has K = 27 and
                # never takes this branch with the tuned S = Sy = 2 packing)
                fuse = cin_total * KHc * KW > 400 or self.ov.get("fused_pool") == "all"
            elif S == 2 and Sy == 2 and MT == 1 and ntot == ntot_pad == NT and not self.ov.get("no_fused_pool2"):
                # the 2x2 block of a pooled pixel is ONE GEMM row: parked in shared memory per 16-channel unit, reduced after a named
                # barrier of the lane quarter's warps (csrc/fcn_conv.cu, kPOOL2) -- conv_down_block_1, whose separate pool pass re-read 1.5 GB
                fuse = True
            elif S >= 2 and S % 2 == 0 and Sy == 1 and MT == 1 and d.RT <= 16 and ntot == ntot_pad == NT and not self.ov.get("no_fused_poolx"):
                # x neighbours = the same 16-channel slice of two column units, which one epilogue warp visits back to back (first unit kept
                # in registers); y neighbours = lanes l, l ^ RT: no shared memory, no barrier (csrc/fcn_conv.cu, kPOOLX)
                fuse = True
        if fuse:
            d.pool_out, d.pool_H, d.pool_W = pool_dst.ptr, pool_dst.H, pool_dst.W
            d.pool_sx, d.pool_sy, d.pool_sn = pool_dst.C, pool_dst.Wp * pool_dst.C, pool_dst.H * pool_dst.Wp * pool_dst.C
            d.pool_padx = pool_dst.pad
        return d, [packed, bias]

    def conv_candidates(self, name):
        """[(modelled cycles, S, Sy, NT, MT)] of a recorded layer, best first (tools/autotune_fcn.py)."""
        sp = self.specs[name]
        if sp["kind"] == "tconv":
            src, cout = sp["src"], sp["wt"].shape[1]
            out = []
            for nt in nt_candidates(4 * cout):
                for mt in MT_OPTIONS:
                    c = layer_cost(src.W, src.H, 4 * cout, [src.C], 1, self.B, nt=nt, mt=mt)
                    if c is not None:
                        out.append((c["clk"], 1, 1, nt, mt))
            return sorted(out)
        w, srcs = sp["w"], sp["srcs"]
        return sorted(self._candidates(srcs[0][0].W, srcs[0][0].H, w.shape[0], [buf.C for buf, _ in srcs], w.shape[3], w.shape[2], sp["cap"]))

    def conv_variant(self, name, cfg):
        """Descriptor of layer `name` under another configuration, reading and writing the plan's own buffers."""
        sp = self.specs[name]
        if sp["kind"] == "tconv":
            return self._tconv_desc(sp["wt"], sp["bt"], sp["src"], sp["dst"], cfg[2], cfg[3])[:2]
        return self._conv_desc(sp["w"], sp["b"], sp["srcs"], sp["dst"], sp["act"], sp["f32_out"], tuple(cfg), sp.get("pool_dst"),
                               sp.get("epi", AM_EPI_PLAIN))

    def _tconv(self, name, wt, bt, src, dst):
        """ConvTranspose2d(k=2,s=2) + BN + GELU as a 1x1 GEMM with N = (sy,sx,co); odd output sizes get the
        bias-only row/column (output_padding, FCN_lecturenet.py:280) from a border fill."""
        cin, cout = wt.shape[0], wt.shape[1]
        forced = self.ov.get("cfg", {}).get(name)
        if forced is None and self.rowrun and "sy" not in self.ov and "mt" not in self.ov and not self.ov.get("no_tuned"):
            forced = tuned_table().get(self.arch, {}).get("B%d_%dx%d" % (self.B, self.H, self.W), {}).get(name)
            if forced is not None and layer_cost(src.W, src.H, 4 * cout, [src.C], 1, self.B, nt=forced[2], mt=forced[3]) is None:
                forced = None                                             # not feasible for this shape: fall back to the model
        if forced is not None:
            NT, MT = forced[2], forced[3]
        else:
            best = None
            cands = [(nt, mt) for nt in nt_candidates(4 * cout) for mt in MT_OPTIONS
                     if layer_cost(src.W, src.H, 4 * cout, [src.C], 1, self.B, nt=nt, mt=mt) is not None]
            if "mt" in self.ov and any(mt == self.ov["mt"] for _, mt in cands):
                cands = [c for c in cands if c[1] == self.ov["mt"]]
            for nt, mt in cands:
                    c = layer_cost(src.W, src.H, 4 * cout, [src.C], 1, self.B, nt=nt, mt=mt)
                    if c is not None and (best is None or c["clk"] < best[0] * 0.97):
                        best = (c["clk"], nt, mt)
            NT, MT = best[1], best[2]
        self.specs[name] = dict(kind="tconv", wt=wt, bt=bt, src=src, dst=dst, op=len(self.ops), cfg=(1, 1, NT, MT))
        d, keep, gelu_b = self._tconv_desc(wt, bt, src, dst, NT, MT)
        self.ops.append(("conv", d))
        self.keep += keep + [gelu_b]
        self.op_flops[len(self.ops) - 1] = 2 * src.H * src.W * 4 * cout * cin
        if dst.H > 2 * src.H or dst.W > 2 * src.W:
            self.ops.append(("border", (dst, 2 * src.H, 2 * src.W, gelu_b)))
        self.flops += 2 * src.H * src.W * 4 * cout * cin

    def _tconv_desc(self, wt, bt, src, dst, NT, MT):
        cin, cout = wt.shape[0], wt.shape[1]
        w = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1).contiguous()      # n = (sy*2+sx)*Cout + co
        packed, bias, ntot, ntot_pad = pack_weights(w, bt.repeat(4), [(src.C, list(range(src.C)))], 1, self.rowrun, NT)
        packed, bias = packed.to(self.device), bias.to(self.device)
        d = ConvDesc()
        d.nseg = 1
        g = d.seg[0]
        g.ptr, g.C, g.Wp, g.Hbuf, g.x_off = src.ptr, src.C, src.Wp, src.H, src.pad
        g.rowrun, g.S, g.run_len, g.KW = int(self.rowrun), 1, src.C, 1
        d.weights, d.bias = packed.data_ptr(), bias.data_ptr()
        d.KH, d.padY, d.in_ystep = 1, 0, 1
        d.nR, d.Hin, d.batch = src.W, src.H, self.B
        d.RT, d.YT = choose_tile(src.W, src.H, 1)
        d.NT, d.Ntot, d.Ntot_pad = NT, ntot, ntot_pad
        d.out, d.out_f32 = dst.ptr, 0
        d.out_H, d.out_W = dst.H, dst.W
        d.out_sx, d.out_sy, d.out_sn = dst.C, dst.Wp * dst.C, dst.H * dst.Wp * dst.C
        d.out_padx, d.out_coff = dst.pad, 0
        d.Cout, d.Sy, d.Sx, d.act = cout, 2, 2, 1
        d.flags = MT_FLAGS[MT]
        gelu_b = (0.5 * bt.double() * (1.0 + torch.erf(bt.double() / math.sqrt(2.0)))).float().to(torch.bfloat16).to(self.device)
        return d, [packed, bias], gelu_b

    # -------------------------------------------------------------------------------------------------
    def _conv_plan(self, i, desc):
        """Prepared launch (tensor maps + tiling) of conv op i, created on first use (needs the device)."""
        h = self._cplans.get(i)
        if h is None:
            h = self.lib.am_conv_plan_create(ctypes.byref(desc))
            if not h:
                raise _lib.AccessMathB200Error("am_conv_plan_create failed for op %d" % i)
            self._cplans[i] = h
        return h

    def conv_plan_info(self, i):
        """[MT, resident weights, accumulator stages, A stages, B stages, grid, smem bytes, work items] of conv op i."""
        info = (ctypes.c_int * 8)()
        _lib.check(self.lib.am_conv_plan_info(self._conv_plan(i, self.ops[i][1]), info), "am_conv_plan_info")
        return list(info)

    def executed_flops(self, i):
        """FLOPs the tensor cores execute for conv op i per frame: valid GEMM rows x padded N x executed K (Toeplitz taps of the S / Sy
        packing, channel padding, 16-wide K steps), with the half-width edge taps of the CTA-pair kernel counted as halves.  The ratio
        to op_flops[i] (algorithmic) is the padding overhead `--layer-table` reports."""
        d = self.ops[i][1]
        k_per_tap = 0
        for s in range(d.nseg):
            g = d.seg[s]
            length, reps = (g.run_len, 1) if g.rowrun else (g.C, g.KW)
            nck = (length + 63) // 64
            k_per_tap += reps * ((nck - 1) * 64 + ((length - (nck - 1) * 64) + 15) // 16 * 16)
        taps = float(d.KH)
        if (d.flags & AM_CONV_CTA_PAIR) and d.in_ystep == 2 and d.Sy == 2 and d.KH >= 4 and d.Ntot_pad == d.NT == d.Ntot and d.NT % 32 == 0 \
                and d.NT >= 128:
            taps -= 1.0                                                   # two edge taps at N / 2 (csrc/fcn_conv.cu: edge_half)
        return 2.0 * d.Hin * d.nR * d.Ntot_pad * k_per_tap * taps

    def __del__(self):
        try:
            for h in self._cplans.values():
                self.lib.am_conv_plan_destroy(h)
            self._cplans = {}
        except Exception:
            pass

    @property
    def launches_per_run(self):
        """Kernel launches of the NEXT run(): the border fills (constants, outside what the transposed convs write) only run once."""
        n_border = sum(1 for k, _ in self.ops if k == "border")
        return 1 + len(self.ops) - (n_border if self._borders_filled else 0)

    def run(self, stream, want_others=False, threshold=128, timing=None, frames=None, want_logits=True):
        """frames (uint8 BGR [B][H][W][3] device tensor; default self.frames) -> self.logits / self.bits (and text_logit /
        rec when asked).  The plan is shared by every extractor of one (net, batch, size): callers that own their input
        buffers pass them here instead of rebinding self.frames.
        want_logits=False (the fused pipelines): the fp32 logits never reach HBM, only the bit-packed mask does.
        timing: optional list; gets (op_index, start_event, end_event) per conv GEMM launch (CUDA events on the
        launching stream, which must be torch's current stream)."""
        lib, B, H, W = _lib.lib(), self.B, self.H, self.W
        st = ctypes.c_void_p(stream)
        chk = _lib.check
        frames = self.frames if frames is None else frames
        assert frames.is_cuda and frames.dtype == torch.uint8 and tuple(frames.shape) == (B, H, W, 3) and frames.is_contiguous()
        chk(lib.am_fcn_prep_input(frames.data_ptr(), B, H, W, self.x0.ptr, self.x0.C, self.x0.pad, st), "am_fcn_prep_input")
        # per-call pointers of the two fused epilogues (kernel arguments of the prepared launches)
        vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        hp = self._conv_plan(self.heads_op, self.ops[self.heads_op][1])
        chk(lib.am_conv_plan_bind(hp, AM_BIND_FRAMES, vp(frames)), "am_conv_plan_bind")
        chk(lib.am_conv_plan_bind(hp, AM_BIND_TEXT_OUT, vp(self.text_logit if want_others else None)), "am_conv_plan_bind")
        chk(lib.am_conv_plan_bind(hp, AM_BIND_REC_OUT, vp(self.rec if want_others else None)), "am_conv_plan_bind")
        if self.fused_threshold:
            op = self._conv_plan(self.out_op, self.ops[self.out_op][1])
            chk(lib.am_conv_plan_bind(op, AM_BIND_OUT, vp(self.logits if want_logits else None)), "am_conv_plan_bind")
            chk(lib.am_conv_plan_bind(op, AM_BIND_THRESHOLD, ctypes.c_void_p(int(threshold))), "am_conv_plan_bind")
        for i, (kind, a) in enumerate(self.ops):
            if kind == "conv":
                if timing is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                chk(lib.am_conv_plan_launch(self._conv_plan(i, a), st), "am_conv_plan_launch")
                if timing is not None:
                    e1.record()
                    timing.append((i, e0, e1))
            elif kind == "pool":
                src, dst = a
                chk(lib.am_fcn_maxpool2(src.ptr, B, src.H, src.W, src.C, src.pad, dst.ptr, dst.pad, st), "am_fcn_maxpool2")
            elif kind == "border":
                if self._borders_filled:                 # gelu(bias) constants in the rows / columns no conv launch writes: filled by the first run
                    continue
                dst, yf, xf, vals = a
                chk(lib.am_fcn_fill_border(dst.ptr, B, dst.H, dst.W, dst.C, dst.pad, yf, xf, vals.data_ptr(), st), "am_fcn_fill_border")
            elif kind == "threshold":
                chk(lib.am_fcn_threshold_pack(self.logits.data_ptr(), B, H, W, int(threshold), self.bits.data_ptr(), st), "am_fcn_threshold_pack")
        self._borders_filled = True


# ----------------------------------------------------------------------------------------------------
class FCN_LectureNet:
    """Drop-in for the reference class (inference only; the training branches are out of scope)."""

    def __init__(self, channels, n_conv_down_1, n_conv_down_2, n_conv_down_3, n_conv_down_4, n_conv_down_5, mid_block,
                 n_upsample_5, n_conv_up_5, n_upsample_4, n_conv_up_4, n_upsample_3, n_conv_up_3,
                 n_upsample_2, n_conv_up_2, n_upsample_1, n_conv_up_1, kernel_size,
                 n_pmaps_1, n_pmaps_2, pixel_kernel_size, reconstruction_mode):
        if channels != 3 or reconstruction_mode:
            raise NotImplementedError("only the 3-channel binarizer branch (reconstruction_mode=False) is on the hot path")
        self.params = _Params(channels, [n_conv_down_1, n_conv_down_2, n_conv_down_3, n_conv_down_4, n_conv_down_5], mid_block,
                              [n_upsample_1, n_upsample_2, n_upsample_3, n_upsample_4, n_upsample_5],
                              [n_conv_up_1, n_conv_up_2, n_conv_up_3, n_conv_up_4, n_conv_up_5], kernel_size,
                              n_pmaps_1, n_pmaps_2, pixel_kernel_size)
        self.reconstruction_mode = reconstruction_mode
        self._plans = {}
        self._device = None
        self.rowrun = True
        self.plan_overrides = None

    # ---- reference-compatible plumbing ---------------------------------------------------------------
    @staticmethod
    def CreateFromConfig(config, in_channels, reconstruction_mode):
        g = config.get
        return FCN_LectureNet(
            in_channels,
            g("FCN_BINARIZER_NET_DOWN_CONV_FILTERS_1", 16), g("FCN_BINARIZER_NET_DOWN_CONV_FILTERS_2", 32),
            g("FCN_BINARIZER_NET_DOWN_CONV_FILTERS_3", 64), g("FCN_BINARIZER_NET_DOWN_CONV_FILTERS_4", 128),
            g("FCN_BINARIZER_NET_DOWN_CONV_FILTERS_5", 256), g("FCN_BINARIZER_NET_MIDDLE_CONV_FILTERS_MIDDLE", 512),
            g("FCN_BINARIZER_NET_UPSAMPLE_FILTERS_5", 256), g("FCN_BINARIZER_NET_UP_CONV_FILTERS_5", 256),
            g("FCN_BINARIZER_NET_UPSAMPLE_FILTERS_4", 128), g("FCN_BINARIZER_NET_UP_CONV_FILTERS_4", 128),
            g("FCN_BINARIZER_NET_UPSAMPLE_FILTERS_3", 64), g("FCN_BINARIZER_NET_UP_CONV_FILTERS_3", 64),
            g("FCN_BINARIZER_NET_UPSAMPLE_FILTERS_2", 32), g("FCN_BINARIZER_NET_UP_CONV_FILTERS_2", 32),
            g("FCN_BINARIZER_NET_UPSAMPLE_FILTERS_1", 16), g("FCN_BINARIZER_NET_UP_CONV_FILTERS_1", 16),
            g("FCN_BINARIZER_NET_KERNEL_SIZE", 3),
            g("FCN_BINARIZER_NET_PIXEL_FEATURES_1", 32), g("FCN_BINARIZER_NET_PIXEL_FEATURES_2", 16),
            g("FCN_BINARIZER_NET_PIXEL_KERNEL_SIZE", 3), reconstruction_mode)

    def state_dict(self):
        return self.params.state_dict()

    def load_state_dict(self, sd, strict=True):
        self._plans.clear()
        return self.params.load_state_dict(sd, strict=strict)

    def parameters(self):
        return self.params.parameters()

    def eval(self):
        self.params.eval()
        return self

    def cuda(self, device=None):
        self._device = torch.device("cuda:%d" % (torch.cuda.current_device() if device is None else int(device)))
        return self

    def to(self, device):
        return self.cuda(torch.device(device).index)

    def __call__(self, x0):
        return self.forward(x0)

    # ---- device plans ------------------------------------------------------------------------------------
    def plan(self, B, H, W):
        if self._device is None:
            _lib.lib()      # raises loudly without a GPU: there is no CPU path
            self.cuda()
        key = (B, H, W, self.rowrun, repr(self.plan_overrides))
        if key not in self._plans:
            with torch.cuda.device(self._device):
                self._plans[key] = FCNPlan(self.params, B, H, W, self._device, self.rowrun, self.plan_overrides)
        return self._plans[key]

    def binarize_frames(self, frames_bgr, want_others=False, threshold=128):
        """uint8 BGR frames (B,H,W,3) (numpy, pinned/CPU tensor or CUDA tensor) -> the plan holding
        logits (B,H,W) fp32 and the bit-packed INK mask (B,H,WPR), all on the device.  The fused fast path."""
        t = torch.from_numpy(frames_bgr) if isinstance(frames_bgr, np.ndarray) else frames_bgr
        B, H, W, _ = t.shape
        if W * H > 2500000:                                              # :434-437 / :481-494 on the device (csrc/resize.cu)
            from .large_frames import LargeView
            ad = self.large_adapter(B, H, W)
            plan = self.plan(B, ad.fcn_height, ad.fcn_width)
            with torch.cuda.device(self._device):
                st = torch.cuda.current_stream().cuda_stream
                ad.frames.copy_(t, non_blocking=True)
                ad.downscale(plan.frames, st)
                plan.run(st, want_others, threshold)
                ad.upscale_bits(plan.bits, st)
            return LargeView(plan, ad)
        plan = self.plan(B, H, W)
        plan.frames.copy_(t, non_blocking=True)
        plan.run(torch.cuda.current_stream().cuda_stream, want_others, threshold)
        return plan

    def large_adapter(self, B, H, W):
        """Device-side 2.5 MP guard for (B, H, W) frames (LANCZOS halving in, NEAREST mask resize out)."""
        from .large_frames import LargeFrameAdapter
        if self._device is None:
            _lib.lib()
            self.cuda()
        key = ("large", B, H, W)
        if key not in self._plans:
            with torch.cuda.device(self._device):
                self._plans[key] = LargeFrameAdapter(B, H, W, self._device)
        return self._plans[key]

    def masks_from_plan(self, plan, f, want_others=True, threshold=128):
        """Reference-format host outputs of frame f: ink mask uint8 (ink = 255, i.e. after `255 - binary`),
        text mask uint8 0/255, reconstructed image uint8 BGR  (FCN_lecturenet.py:461-479)."""
        from .cc_engine import _p, _stream
        ink = torch.empty((1, plan.H, plan.W), dtype=torch.uint8, device=plan.device)
        _lib.check(plan.lib.am_unpack_mask_u8(_p(plan.bits[f:f + 1]), plan.W, plan.H, 1, _p(ink), _stream()), "am_unpack_mask_u8")
        binary = ink[0].cpu().numpy()
        if not want_others:
            return binary, None, None
        t = (torch.sigmoid(plan.text_logit[f]).cpu().numpy() * 255).astype(np.uint8)
        text_mask = np.where(t >= threshold, 255, 0).astype(np.uint8)
        rec = plan.rec[f].cpu().numpy() * 0.5 + 0.5
        rec_img = np.clip(rec[:, :, ::-1] * 255, 0, 255).astype(np.uint8)
        if text_mask.shape != (plan.H, plan.W):                          # > 2.5 MP frame: the diagnostic outputs follow :487-494
            import cv2
            text_mask = cv2.resize(text_mask, (plan.W, plan.H), interpolation=cv2.INTER_NEAREST)
            rec_img = cv2.resize(rec_img, (plan.W, plan.H), interpolation=cv2.INTER_NEAREST)
        return binary, text_mask, rec_img

    @staticmethod
    def prepare_image(PIL_image):
        """(1,3,H,W) fp32 in [-1,1]  (:607-618)"""
        a = np.asarray(PIL_image.convert("RGB"), dtype=np.uint8)
        t = torch.from_numpy(a.copy()).permute(2, 0, 1).float().div(255.0)
        return ((t - 0.5) / 0.5).unsqueeze(0)

    def forward(self, x0):
        """x0: (N,3,H,W) fp32 normalised as prepare_image does -> (output_logit, text_mask_logit, rec_img) CUDA fp32.
        The kernels start from the uint8 frame, so x0 is mapped back to its uint8 pixels (exact for prepare_image output)."""
        x = x0.detach().float().cpu()
        u8 = torch.round((x * 0.5 + 0.5) * 255.0).clamp(0, 255).to(torch.uint8)
        if (((u8.float() / 255.0) - 0.5) / 0.5 - x).abs().max().item() > 1e-6:
            raise ValueError("FCN_LectureNet.forward: the device path starts from uint8 pixels, so x0 must be prepare_image() output "
                             "(values (v/255 - 0.5)/0.5 with integer v); got a tensor that is not representable that way")
        bgr = u8.permute(0, 2, 3, 1).flip(-1).contiguous()
        plan = self.binarize_frames(bgr, want_others=True)
        return (plan.logits.unsqueeze(1).clone(), plan.text_logit.unsqueeze(1).clone(), plan.rec.permute(0, 3, 1, 2).contiguous())

    def binarize(self, PIL_image, return_others=False, force_binary=False, binary_treshold=128, apply_sigmoid=True):
        import cv2
        o_width, o_height = PIL_image.size
        width, height = o_width, o_height
        rgb = np.asarray(PIL_image.convert("RGB"), dtype=np.uint8)
        plan = self.binarize_frames(np.ascontiguousarray(rgb[None, :, :, ::-1]), want_others=return_others)
        if width * height > 2500000:                                     # :434-437 ran on the device (am_lanczos_resize_u8)
            width, height = plan.adapter.fcn_width, plan.adapter.fcn_height
        res = plan.logits[0]
        text = plan.text_logit[0] if return_others else None
        if apply_sigmoid:                                                # :452-454
            res = torch.sigmoid(res)
            text = torch.sigmoid(text) if return_others else None
        binary = (res.cpu().numpy() * 255).astype(np.uint8)              # :461-462
        if force_binary:
            binary[binary >= binary_treshold] = 255
            binary[binary < binary_treshold] = 0
        if return_others:
            text_mask = (text.cpu().numpy() * 255).astype(np.uint8)
            if force_binary:
                text_mask[text_mask >= binary_treshold] = 255
                text_mask[text_mask < binary_treshold] = 0
            rec = plan.rec[0].cpu().numpy() * 0.5 + 0.5                  # from_img_space_to_cv2 (:532-554)
            rec_img = np.clip(rec[:, :, ::-1] * 255, 0, 255).astype(np.uint8)
        if o_width != width:                                             # :481-494
            interp = cv2.INTER_NEAREST if force_binary else cv2.INTER_CUBIC
            binary = cv2.resize(binary, (o_width, o_height), interpolation=interp)
            if return_others:
                text_mask = cv2.resize(text_mask, (o_width, o_height), interpolation=interp)
                rec_img = cv2.resize(rec_img, (o_width, o_height), interpolation=cv2.INTER_NEAREST)
        if return_others:
            return binary, text_mask, rec_img
        return binary
