"""Distributed result checks for the slab-decomposed transform (no CPU FFT involved: float64 direct sums).

A distributed 3-D transform can be numerically perfect and still wrong: a Y row stored into the wrong peer's
slab, or z planes of two source ranks swapped, leaves the energy (Parseval) unchanged. Two checks that see
the PLACEMENT of every element, both run with the data where it is (one process per rank):

* sampled_bins_check  — for a fixed list of bins (kz, ky, kx) every rank forms, in float64, its z planes' share
  of the direct DFT sum  sum_{z,y,x} v[z,y,x] exp(-+2 pi i (kz z/Z + ky y/Y + kx x/X)); the shares are summed
  over ranks (all_reduce) and the owner of ky compares with the element it holds. Random data: measures accuracy
  and catches any systematic misplacement (each bin depends on every input element and on its position).
* delta_volume_check  — input = one weighted unit impulse per source rank; the exact spectrum is a sum of G
  plane waves, evaluated in float64 for EVERY element of the local output slab: each output element is compared,
  so a single misplaced row / plane / peer slot shows up, and each source rank's contribution has its own
  amplitude and wave vector.

Layout conventions are SlabFFT3D's: input rank g holds v[g*Z/G:(g+1)*Z/G, :, :] as (Z/G, Y, X, 2) float32,
output rank h holds V[:, h*Y/G:(h+1)*Y/G, :] as (Z, Y/G, X, 2).
"""
import math

import torch
import torch.distributed as dist


def _phasor(n, k, idx, sign):
    """exp(sign * 2 pi i * k * idx / n) as complex128, argument reduced in integers first."""
    r = (k * idx) % n
    ang = r.to(torch.float64) * (sign * 2.0 * math.pi / n)
    return torch.complex(torch.cos(ang), torch.sin(ang))


def default_bins(dims, world, count=32, seed=2024):
    """`count` bins, identical on every rank, touching every output owner (ky block) and the block edges."""
    Z, Y, X = dims
    g = torch.Generator().manual_seed(seed)
    yl = Y // world
    bins = []
    for i in range(count):
        owner = i % world
        ky = owner * yl + (0 if i < world else yl - 1 if i < 2 * world else int(torch.randint(0, yl, (1,), generator=g)))
        kz = int(torch.randint(0, Z, (1,), generator=g))
        kx = int(torch.randint(0, X, (1,), generator=g))
        bins.append((kz, ky, kx))
    bins[0] = (0, 0, 0)
    bins[-1] = (Z - 1, Y - 1, X - 1)
    return bins


def sampled_bins_check(x_local, out_local, dims, rank, world, group=None, bins=None, inverse=False):
    """max over the sampled bins of |got - want| / rms(|want|) (a float, the same on every rank).
    x_local: (Z/G, Y, X, 2) input of this rank; out_local: (Z, Y/G, X, 2) result slab of this rank."""
    Z, Y, X = dims
    zl, yl = Z // world, Y // world
    bins = bins or default_bins(dims, world)
    dev = x_local.device
    sign = 1.0 if inverse else -1.0
    kz = torch.tensor([b[0] for b in bins], device=dev)
    ky = torch.tensor([b[1] for b in bins], device=dev)
    kx = torch.tensor([b[2] for b in bins], device=dev)
    nb = len(bins)
    xs = torch.arange(X, device=dev)
    ys = torch.arange(Y, device=dev)
    zs = torch.arange(rank * zl, (rank + 1) * zl, device=dev)
    ex = _phasor(X, kx[:, None], xs[None, :], sign)        # (nb, X)
    ey = _phasor(Y, ky[:, None], ys[None, :], sign)        # (nb, Y)
    ez = _phasor(Z, kz[:, None], zs[None, :], sign)        # (nb, zl)
    part = torch.zeros(nb, dtype=torch.complex128, device=dev)
    # plane by plane keeps the float64 temporaries at (Y, X) + (Y, nb)
    for z in range(zl):
        v = torch.view_as_complex(x_local[z].to(torch.float64).contiguous())   # (Y, X)
        t = v @ ex.transpose(0, 1)                                            # (Y, nb)
        part += (t * ey.transpose(0, 1)).sum(0) * ez[:, z]
    if inverse:
        part = part / float(Z * Y * X)
    want = torch.view_as_real(part).contiguous()
    got = torch.zeros((nb, 2), dtype=torch.float64, device=dev)
    for i, (bz, by, bx) in enumerate(bins):
        if by // yl == rank:
            got[i] = out_local[bz, by % yl, bx].to(torch.float64)
    if world > 1:
        dist.all_reduce(want, group=group)
        dist.all_reduce(got, group=group)
    d = torch.view_as_complex(got) - torch.view_as_complex(want)
    rms = torch.view_as_complex(want).abs().pow(2).mean().sqrt()
    return float(d.abs().max() / rms)


def delta_input(dims, rank, world, device, dtype=torch.float32):
    """This rank's planes of the impulse volume: one impulse per source rank g at
    (z, y, x) = (g*Z/G + (3g+1) % (Z/G), (5g+3) % Y, (7g+5) % X) with amplitude g + 1 + i(g + 2)/2."""
    Z, Y, X = dims
    zl = Z // world
    x = torch.zeros((zl, Y, X, 2), device=device, dtype=dtype)
    z0, y0, x0, a = _delta_of(dims, world, rank)
    x[z0 - rank * zl, y0, x0, 0] = a.real
    x[z0 - rank * zl, y0, x0, 1] = a.imag
    return x


def _delta_of(dims, world, g):
    Z, Y, X = dims
    zl = Z // world
    return g * zl + (3 * g + 1) % zl, (5 * g + 3) % Y, (7 * g + 5) % X, complex(g + 1.0, (g + 2.0) / 2.0)


def delta_volume_check(out_local, dims, rank, world, group=None, inverse=False):
    """max |got - want| over EVERY element of every rank's slab for the delta_input() volume, divided by the largest
    |want| (a float, the same on every rank)."""
    Z, Y, X = dims
    yl = Y // world
    dev = out_local.device
    sign = 1.0 if inverse else -1.0
    ys = torch.arange(rank * yl, (rank + 1) * yl, device=dev)
    xs = torch.arange(X, device=dev)
    stat = torch.zeros(2, dtype=torch.float64, device=dev)
    zc = max(1, min(Z, (1 << 22) // max(1, yl * X)))      # float64 temporaries of ~4 M complex points per chunk
    for z0c in range(0, Z, zc):
        zs = torch.arange(z0c, min(Z, z0c + zc), device=dev)
        want = torch.zeros((zs.numel(), yl, X), dtype=torch.complex128, device=dev)
        for g in range(world):
            z0, y0, x0, a = _delta_of(dims, world, g)
            pz = _phasor(Z, zs, torch.tensor(z0, device=dev), sign)
            py = _phasor(Y, ys, torch.tensor(y0, device=dev), sign)
            px = _phasor(X, xs, torch.tensor(x0, device=dev), sign)
            want += a * pz[:, None, None] * py[None, :, None] * px[None, None, :]
        if inverse:
            want = want / float(Z * Y * X)
        got = torch.view_as_complex(out_local[z0c:z0c + zs.numel()].to(torch.float64).contiguous())
        stat = torch.maximum(stat, torch.stack([(got - want).abs().max(), want.abs().max()]))
    if world > 1:
        dist.all_reduce(stat, op=dist.ReduceOp.MAX, group=group)
    return float(stat[0] / stat[1])
