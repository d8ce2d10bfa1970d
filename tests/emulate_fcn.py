"""Test infrastructure: a torch/CPU emulation of what csrc/fcn_conv.cu + fcn_misc.cu compute from an FCNPlan
(same packed weights, same descriptors, same buffer layout, TMA out-of-bounds zero fill, K-step masking, epilogue
scatter).  It checks the HOST logic (weight packing, S-packing, K segments, tile/epilogue index math) on the CPU
tier; the GPU tier then only has the PTX itself left to prove."""
import math

import torch


def _flat_of(plan):
    m = {}
    for b in [plan.x0, plan.mid, plan.diff, plan.px1, plan.px2] + plan.d + plan.p + plan.t + plan.u:
        m[b.ptr] = b.t
    for t in plan.keep + [plan.logits, plan.text_logit, plan.rec, plan.frames]:
        m[t.data_ptr()] = t.view(-1)
    m[plan.bits.data_ptr()] = plan.bits
    return m


def emulate_conv(d, mem):
    B, Hin, nR, KH = d.batch, d.Hin, d.nR, d.KH
    Wmat = mem[d.weights].float().view(-1, KH, d.Ntot_pad, 64)           # [chunk][dy][n][64]
    bias = mem[d.bias].float()
    acc = torch.zeros((B, Hin, nR, d.Ntot_pad), dtype=torch.float32)
    chunk = 0
    n_i = torch.arange(B).view(B, 1, 1, 1)
    y_i = torch.arange(Hin).view(1, Hin, 1, 1)
    r_i = torch.arange(nR).view(1, 1, nR, 1)
    kk = torch.arange(64).view(1, 1, 1, 64)
    for s in range(d.nseg):
        g = d.seg[s]
        flat = mem[g.ptr].float()
        base = g.x_off * g.C
        if g.rowrun:
            nkx, nck, dim0 = 1, (g.run_len + 63) // 64, g.run_len
            klast = ((g.run_len - (nck - 1) * 64) + 15) // 16 * 16
        else:
            nkx, nck, dim0 = g.KW, (g.C + 63) // 64, g.C
            klast = ((g.C - (nck - 1) * 64) + 15) // 16 * 16
        for kx in range(nkx):
            for ck in range(nck):
                kvalid = klast if ck == nck - 1 else 64
                ystep = 2 if d.in_ystep == 2 else 1
                for dy in range(KH):
                    yy = y_i * ystep + dy - d.padY
                    k = ck * 64 + kk
                    if g.rowrun:
                        idx = base + ((n_i * g.Hbuf + yy) * g.Wp) * g.C + r_i * g.S * g.C + k
                        ok = (yy >= 0) & (yy < g.Hbuf) & (k < dim0)
                    else:
                        x = r_i + kx
                        idx = base + ((n_i * g.Hbuf + yy) * g.Wp + x) * g.C + k
                        ok = (yy >= 0) & (yy < g.Hbuf) & (k < dim0) & (x < g.Wp)
                    ok = ok & (kk < kvalid)
                    idx = torch.where(ok, idx, torch.zeros_like(idx))
                    a = torch.where(ok, flat[idx.reshape(-1)].view(idx.shape), torch.zeros(()))
                    acc += torch.einsum("byrk,nk->byrn", a, Wmat[chunk, dy])
                chunk += 1
    x = acc + bias.view(1, 1, 1, -1)
    if d.epi_mode:                                                         # fused epilogues: gather the fp32 planes the kernel keeps in registers
        planes = torch.zeros((B, d.out_H, d.out_W, d.Cout), dtype=torch.float32)
        for j in range(d.Ntot):
            grp, co = j // d.Cout, j % d.Cout
            sy, sx = grp // d.Sx, grp % d.Sx
            oy, ox = d.Sy * torch.arange(Hin) + sy, d.Sx * torch.arange(nR) + sx
            vy, vx = oy < d.out_H, ox < d.out_W
            planes[:, oy[vy].view(-1, 1), ox[vx].view(1, -1), co] = x[:, vy][:, :, vx][..., j]
        if d.epi_mode == 1:                                                # AM_EPI_HEADS
            fr = mem[d.frames].view(B, d.out_H, d.out_W, 3)
            rgb = _norm(fr.flip(-1))
            text, rec = planes[..., 0], torch.tanh(planes[..., 1:4])
            diff = ((rgb - rec) * torch.sigmoid(text).unsqueeze(-1)).to(torch.bfloat16)
            dbuf = mem[d.diff_out][:B * d.out_H * (d.out_W + 2 * d.diff_pad) * d.diff_C].view(B, d.out_H, d.out_W + 2 * d.diff_pad, d.diff_C)
            dbuf[:, :, d.diff_pad:d.diff_pad + d.out_W, :3] = diff
            dbuf[:, :, d.diff_pad:d.diff_pad + d.out_W, 3:] = 0
            if d.text_out:
                mem[d.text_out].view(B, d.out_H, d.out_W)[:] = text
            if d.rec_out:
                mem[d.rec_out].view(B, d.out_H, d.out_W, 3)[:] = rec
        else:                                                              # AM_EPI_THRESHOLD
            assert d.Cout == 1 and d.Sx % 16 == 0 and d.out_W % 16 == 0
            z = planes[..., 0]
            ink = ((torch.sigmoid(z).numpy() * 255).astype("uint8") < d.threshold)
            import numpy as np
            packed = np.packbits(np.pad(ink, ((0, 0), (0, 0), (0, d.bits_wpr * 32 - d.out_W))), axis=-1, bitorder="little")
            mem[d.bits_out][:] = torch.from_numpy(packed.view("<i4").reshape(B, d.out_H, d.bits_wpr))
            if d.out:
                mem[d.out].view(B, d.out_H, d.out_W)[:] = z
        return
    out = mem[d.out]
    if d.act == 1:
        x = 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
    n = torch.arange(d.Ntot)
    grp, co = n // d.Cout, n % d.Cout
    sy, sx = grp // d.Sx, grp % d.Sx
    for j in range(d.Ntot):
        oy = d.Sy * torch.arange(Hin) + int(sy[j])
        ox = d.Sx * torch.arange(nR) + int(sx[j])
        vy, vx = oy < d.out_H, ox < d.out_W
        off = (torch.arange(B).view(B, 1, 1) * d.out_sn + oy[vy].view(1, -1, 1) * d.out_sy +
               (ox[vx].view(1, 1, -1) + d.out_padx) * d.out_sx + d.out_coff + int(co[j]))
        vals = x[:, vy][:, :, vx][..., j]
        out[off.reshape(-1)] = vals.reshape(-1).to(out.dtype)
    if d.pool_out:                                                         # fused MaxPool2d(2) of the bf16 output
        pout = mem[d.pool_out]
        if d.Sx == 2 and d.Sy == 2:                                        # kPOOL2: the 2x2 block is the four column groups of one GEMM row
            assert d.Ntot == 4 * d.Cout == d.NT
            v = x[:, :d.pool_H, :d.pool_W, :d.Ntot].to(torch.bfloat16).float().reshape(B, d.pool_H, d.pool_W, 4, d.Cout).amax(dim=3)
        elif d.Sx >= 2:                                                    # kPOOLX: x pairs in column units, y pairs in GEMM rows
            assert d.Sx % 2 == 0 and d.Sy == 1 and d.RT <= 16 and d.Ntot == d.Sx * d.Cout == d.NT
            v = x[..., :d.Ntot].to(torch.bfloat16).float().reshape(B, Hin, nR * d.Sx, d.Cout)[:, :d.pool_H * 2, :d.pool_W * 2]
            v = v.reshape(B, d.pool_H, 2, d.pool_W, 2, d.Cout).amax(dim=(2, 4))
        else:
            assert d.Sx == 1 and d.Sy == 1 and d.RT <= 16
            v = x[:, :d.pool_H * 2, :d.pool_W * 2, :d.Cout].to(torch.bfloat16).float()
            v = v.reshape(B, d.pool_H, 2, d.pool_W, 2, d.Cout).amax(dim=(2, 4))
        off = (torch.arange(B).view(B, 1, 1, 1) * d.pool_sn + torch.arange(d.pool_H).view(1, -1, 1, 1) * d.pool_sy +
               (torch.arange(d.pool_W).view(1, 1, -1, 1) + d.pool_padx) * d.pool_sx + torch.arange(d.Cout).view(1, 1, 1, -1))
        pout[off.reshape(-1)] = v.reshape(-1).to(pout.dtype)


def _norm(u8):
    return (u8.float() / 255.0 - 0.5) / 0.5


def emulate_plan(plan, frames_bgr):
    """Run the whole plan on the CPU; returns (logits, text_logit, rec, ink)."""
    mem = _flat_of(plan)
    B, H, W = plan.B, plan.H, plan.W
    fr = torch.as_tensor(frames_bgr)
    rgb = _norm(fr.flip(-1))
    plan.x0.view()[..., :3] = rgb.to(torch.bfloat16)
    plan.frames.copy_(fr)
    for kind, a in plan.ops:
        if kind == "conv" and a.epi_mode == 1:                             # what FCNPlan.run binds per call
            a.text_out, a.rec_out = plan.text_logit.data_ptr(), plan.rec.data_ptr()
        if kind == "conv":
            emulate_conv(a, mem)
        elif kind == "pool":
            src, dst = a
            v = src.view().float()
            Ho, Wo = src.H // 2, src.W // 2
            v = v[:, :Ho * 2, :Wo * 2].reshape(B, Ho, 2, Wo, 2, src.C).amax(dim=(2, 4))
            dst.view()[:] = v.to(torch.bfloat16)
        elif kind == "border":
            dst, yf, xf, vals = a
            v = dst.view()
            v[:, yf:, :, :] = vals
            v[:, :, xf:, :] = vals
        elif kind == "threshold":
            pass
    text, rec = plan.text_logit.clone(), plan.rec.clone()
    logits = plan.logits.clone()
    u8 = (torch.sigmoid(logits).numpy() * 255).astype("uint8")
    ink = (u8 < 128)
    return logits, text, rec, ink
