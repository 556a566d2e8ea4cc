"""Fused SyncBatchNorm exchange (csrc/syncbn.cu) on ONE GPU: `world` ranks are emulated by `world` CUDA streams, each with its own
exchange buffer (ordinary device memory; on the multi-GPU path the same pointers are NVLink-mapped peer buffers from
torch.distributed._symmetric_memory) and its own per-slab partial sums.  The kernels of all ranks run concurrently and wait for each
other's epoch flags exactly as they do across GPUs, so this exercises the whole protocol -- stores to every peer, release / acquire
flags, rank-ordered summation, slot reuse with the next epoch -- through the C ABI.

Checked against a float64 restatement of nn.SyncBatchNorm's arithmetic (statistics over all ranks' rows; torch semantics of
SURVEY.md App. D): mean / invstd / scale / shift and the running statistics to 1e-6, the backward sums to 1e-6, and every rank's
results BIT-identical to rank 0's (the property DDP replicas rely on).
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


class _Ring:
    """`world` exchange buffers + the device pointer table, laid out like parallel.SyncExchange"""

    def __init__(self, lib, world, cmax, max_slots=4):
        self.world, self.cmax = world, cmax
        self.slot_floats = lib.segb200_syncbn_slot_floats(world, cmax)
        self.slot_flags = lib.segb200_syncbn_slot_flags(world)
        self.flags_base = max_slots * self.slot_floats
        n = max_slots * (self.slot_floats + self.slot_flags)
        self.bufs = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(world)]
        self.peers = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device="cuda")
        self.epoch = torch.zeros(1, dtype=torch.int32, device="cuda")

    def slot(self, s):
        return s * self.slot_floats, self.flags_base + s * self.slot_flags


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("c,slabs", [(64, 37), (728, 5), (2048, 130)])
def test_fused_syncbn_exchange_protocol(world, c, slabs):
    from segmentron_b200 import lib as L
    lib = L.load()
    cmax = 2048
    ring = _Ring(lib, world, cmax)
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator().manual_seed(c * 10 + world)
    rows = 4000.0                                              # rows per rank
    # per-rank per-slab partial sums of x and x^2 (consistent: sum x^2 >= (sum x)^2 / rows)
    part = []
    for r in range(world):
        s = torch.randn(slabs, c, generator=g) * 3.0 + 0.5 * (r + 1)
        q = s * s / (rows / slabs) + torch.rand(slabs, c, generator=g) * 50.0 + 1.0
        part.append(torch.stack([s, q], 1).contiguous().cuda())                    # [slabs][2][c]
    gamma = (0.5 + torch.rand(c, generator=g)).cuda()
    beta = torch.randn(c, generator=g).cuda()
    rm0, rv0 = torch.randn(c, generator=g).cuda(), (0.5 + torch.rand(c, generator=g)).cuda()
    mom, eps = 0.1, 1e-3
    out = [dict(mean=torch.empty(c, device="cuda"), invstd=torch.empty(c, device="cuda"), scale=torch.empty(c, device="cuda"),
                shift=torch.empty(c, device="cuda"), rm=rm0.clone(), rv=rv0.clone(), sums=torch.empty(2, c, device="cuda"),
                dgamma=torch.zeros(c, device="cuda"), dbeta=torch.zeros(c, device="cuda")) for _ in range(world)]
    count = rows * world
    torch.cuda.synchronize()

    def launch_fwd(slot):
        d_off, f_off = ring.slot(slot)
        L.check(lib.segb200_counter_add(_ptr(ring.epoch), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "counter_add")
        torch.cuda.synchronize()
        for r in range(world):
            o = out[r]
            L.check(lib.segb200_bn_finalize_sync(_ptr(part[r]), slabs, c, count, _ptr(gamma), _ptr(beta), _ptr(o["rm"]), _ptr(o["rv"]), mom,
                                                 eps, _ptr(o["mean"]), _ptr(o["invstd"]), _ptr(o["scale"]), _ptr(o["shift"]),
                                                 _ptr(ring.peers), world, r, cmax, d_off, f_off, _ptr(ring.epoch),
                                                 C.c_void_p(streams[r].cuda_stream)), "bn_finalize_sync")
        torch.cuda.synchronize()

    launch_fwd(0)
    tot = sum(p.double().sum(0) for p in part)                                      # [2][c] over all ranks and slabs
    # the kernel exchanges fp32 per-rank sums: restate that rounding
    tot32 = sum(p.double().sum(0).float().double() for p in part)
    m = tot32[0] / count
    var = (tot32[1] / count - m * m).clamp(min=0)
    inv = 1.0 / torch.sqrt(var + eps)
    for r in range(world):
        o = out[r]
        assert torch.allclose(o["mean"].double(), m, rtol=1e-6, atol=1e-6)
        assert torch.allclose(o["invstd"].double(), inv, rtol=2e-6)
        assert torch.allclose(o["scale"].double(), gamma.double() * inv, rtol=2e-6)
        assert torch.allclose(o["shift"].double(), beta.double() - m * gamma.double() * inv, rtol=1e-5, atol=1e-5)
        assert torch.allclose(o["rm"].double(), (1 - mom) * rm0.double() + mom * m, rtol=1e-6, atol=1e-6)
        assert torch.allclose(o["rv"].double(), (1 - mom) * rv0.double() + mom * var * count / (count - 1), rtol=1e-5)
        for k in ("mean", "invstd", "scale", "shift", "rm", "rv"):
            assert torch.equal(o[k], out[0][k]), f"rank {r} differs from rank 0 in {k}"
    assert float((tot - tot32).abs().max() / tot.abs().max()) < 1e-6

    # ---- same slot again with the next epoch (a training step later), then the backward exchange on another slot ----
    for p in part:
        p.mul_(1.25)
    launch_fwd(0)
    m2 = sum(p.double().sum(0).float().double() for p in part)[0] / count
    for r in range(world):
        assert torch.allclose(out[r]["mean"].double(), m2, rtol=1e-6, atol=1e-6), "stale data / flag from the previous epoch"

    d_off, f_off = ring.slot(1)
    bpart = [torch.randn(slabs, 2, c, generator=g).cuda() for _ in range(world)]
    mean, invstd = out[0]["mean"].clone(), out[0]["invstd"].clone()
    for r in range(world):
        o = out[r]
        L.check(lib.segb200_bn_bwd_finalize_sync(_ptr(bpart[r]), slabs, c, _ptr(mean), _ptr(invstd), _ptr(o["sums"]), _ptr(o["dgamma"]),
                                                 _ptr(o["dbeta"]), _ptr(ring.peers), world, r, cmax, d_off, f_off, _ptr(ring.epoch),
                                                 C.c_void_p(streams[r].cuda_stream)), "bn_bwd_finalize_sync")
    torch.cuda.synchronize()
    loc = []
    for r in range(world):
        p = bpart[r].double().sum(0)
        loc.append(torch.stack([p[0], invstd.double() * (p[1] - mean.double() * p[0])]))
    tot_b = sum(x.float().double() for x in loc)
    for r in range(world):
        o = out[r]
        assert torch.allclose(o["sums"].double(), tot_b, rtol=1e-5, atol=1e-4)
        assert torch.equal(o["sums"], out[0]["sums"])
        assert torch.allclose(o["dgamma"].double(), loc[r][1], rtol=1e-5, atol=1e-4)      # LOCAL sums: DDP averages parameter gradients
        assert torch.allclose(o["dbeta"].double(), loc[r][0], rtol=1e-5, atol=1e-4)


def test_fused_syncbn_world1_equals_plain_finalize():
    """world == 1: the fused kernel must reproduce segb200_bn_finalize exactly (same fixed-order fp64 reduction is NOT promised -- the
    slab order differs -- so: 1e-6)."""
    from segmentron_b200 import lib as L
    lib = L.load()
    c, slabs, rows = 256, 77, 12345.0
    ring = _Ring(lib, 1, 2048)
    g = torch.Generator().manual_seed(5)
    part = torch.randn(slabs, 2, c, generator=g)
    part[:, 1] = part[:, 0] ** 2 + 10.0
    part = part.cuda()
    gamma, beta = torch.rand(c, generator=g).cuda(), torch.randn(c, generator=g).cuda()
    res = []
    for fused in (False, True):
        o = [torch.empty(c, device="cuda") for _ in range(4)]
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if fused:
            L.check(lib.segb200_counter_add(_ptr(ring.epoch), 1, s), "counter_add")
            d_off, f_off = ring.slot(0)
            L.check(lib.segb200_bn_finalize_sync(_ptr(part), slabs, c, rows, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), 0.1, 1e-5, _ptr(o[0]),
                                                 _ptr(o[1]), _ptr(o[2]), _ptr(o[3]), _ptr(ring.peers), 1, 0, 2048, d_off, f_off,
                                                 _ptr(ring.epoch), s), "bn_finalize_sync")
        else:
            L.check(lib.segb200_bn_finalize(_ptr(part), slabs, c, rows, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), 0.1, 1e-5, _ptr(o[0]),
                                            _ptr(o[1]), _ptr(o[2]), _ptr(o[3]), s), "bn_finalize")
        torch.cuda.synchronize()
        res.append(o + [rm, rv])
    for a, b in zip(*res):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-6)
