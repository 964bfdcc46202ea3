"""Padded-flat (PF) activation layout helpers (torch, host-side plumbing for tests and tools).

row(t, ph, b, y, x) = ((t * P + ph) * B + b) * (H + 2) * (W + 2) + (y + 1) * (W + 2) + (x + 1); see
include/crossloc_b200.h.  The kernels never call these: they exist so that tests can feed single operators.
"""
import torch


def to_pf(x, phases=1, terms=2):
    """NCHW fp32 -> fp16 PF matrix [terms * phases * B * (H'+2) * (W'+2)][C]; (H', W') = (H, W) or halves for phases=4."""
    b, c, h, w = x.shape
    x = x.to(torch.float32)
    if phases == 1:
        planes = [x]
        hh, ww = h, w
    else:
        hh, ww = (h + 1) // 2, (w + 1) // 2
        planes = []
        for a in range(2):
            for bb in range(2):
                p = torch.zeros(b, c, hh, ww, dtype=torch.float32, device=x.device)
                sub = x[:, :, a::2, bb::2]
                p[:, :, :sub.size(2), :sub.size(3)] = sub
                planes.append(p)
    out = []
    for t in range(terms):
        for p in planes:
            hi = p.to(torch.float16)
            val = hi if t == 0 else (p - hi.to(torch.float32)).to(torch.float16)
            pad = torch.zeros(b, hh + 2, ww + 2, c, dtype=torch.float16, device=x.device)
            pad[:, 1:-1, 1:-1, :] = val.permute(0, 2, 3, 1)
            out.append(pad.reshape(-1, c))
    return torch.cat(out, 0).contiguous()


def to_pf8(x, hi_scale=4.0, lo_scale=16384.0):
    """NCHW fp32 -> e4m3 PF planes (uint8 view) [2 * B * (H+2) * (W+2)][C]: fp8(a_hi * 2^2), fp8((a - a_hi) * 2^14)."""
    b, c, h, w = x.shape
    x = x.to(torch.float32)
    hi = x.to(torch.float16).to(torch.float32)
    out = []
    for val in (hi * hi_scale, (x - hi) * lo_scale):
        q = val.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)
        pad = torch.zeros(b, h + 2, w + 2, c, dtype=torch.uint8, device=x.device)
        pad[:, 1:-1, 1:-1, :] = q.permute(0, 2, 3, 1)
        out.append(pad.reshape(-1, c))
    return torch.cat(out, 0).contiguous()


def from_pf(buf, batch, h, w, terms=2):
    """fp16 PF matrix (phases = 1) -> NCHW fp32 (hi + lo)."""
    c = buf.size(1)
    rows = batch * (h + 2) * (w + 2)
    x = buf[:rows].to(torch.float32)
    if terms == 2:
        x = x + buf[rows:2 * rows].to(torch.float32)
    return x.reshape(batch, h + 2, w + 2, c)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).contiguous()


def from_pf_phases(buf, batch, h, w, terms=2):
    """fp16 PF matrix with 4 phases at (ceil(h/2), ceil(w/2)) -> NCHW fp32 at (h, w)."""
    c = buf.size(1)
    hh, ww = (h + 1) // 2, (w + 1) // 2
    rows = batch * (hh + 2) * (ww + 2)
    out = torch.zeros(batch, c, h, w, dtype=torch.float32, device=buf.device)
    for ph in range(4):
        x = buf[ph * rows:(ph + 1) * rows].to(torch.float32)
        if terms == 2:
            x = x + buf[(4 + ph) * rows:(5 + ph) * rows].to(torch.float32)
        x = x.reshape(batch, hh + 2, ww + 2, c)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
        a, bb = ph // 2, ph % 2
        sub = out[:, :, a::2, bb::2]
        sub.copy_(x[:, :, :sub.size(2), :sub.size(3)])
    return out


def raw_to_nchw(raw, batch, h, w):
    """fp32 PF raw matrix [B*(H+2)*(W+2)][C] -> NCHW interior."""
    c = raw.size(1)
    return raw.reshape(batch, h + 2, w + 2, c)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).contiguous()
