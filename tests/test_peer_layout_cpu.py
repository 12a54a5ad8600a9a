"""CPU: the partition arithmetic of the NVLink peer-memory all-reduce (csrc/peer.cu) restated in numpy.

The kernels cannot run here (no GPU, and they need >= 2 of them), but everything that depends on the
number of ranks is index arithmetic: which vectors a rank pushes where (`peer_push_kernel` / phase A of
`peer_allreduce_kernel`), which staged copies a rank sums and where the result goes
(`peer_reduce_kernel` / phase C).  This file replays that arithmetic, CTA by CTA and thread by thread,
for world = 1 ... 8 and awkward slice sizes, on real per-rank buffers, and checks
  * every rank ends up with the mean of all ranks' inputs, identical on all ranks,
  * nothing outside [offset, offset + count) is touched, no staging slot is written twice,
  * the staging buffer size exported by the library (ub200_peer_stage_bytes) covers every index used.
The flag protocol itself (fences, epochs) is rank-count independent and is exercised on hardware by
tools/peer_check.py / tests/test_dist_nccl2_gpu.py."""
import ctypes as C

import numpy as np
import pytest

from uniter_b200 import _lib

PUSH_U, RED_U = 8, 2
PUSH_VPC, RED_VPC = 256 * PUSH_U, 256 * RED_U


def _per(nvec, world):
    per = (nvec + world - 1) // world
    return (per + PUSH_VPC - 1) // PUSH_VPC * PUSH_VPC


def _emulate_v2(bufs, lo_vec, nvec, stage_vecs):
    """bufs[r]: float64 [total_vecs] (one number stands for one 16-byte vector)."""
    world = len(bufs)
    per = _per(nvec, world)
    stage = [np.full(stage_vecs, np.nan) for _ in range(world)]
    written = [np.zeros(stage_vecs, dtype=np.int32) for _ in range(world)]
    tid = np.arange(256)
    for rank in range(world):                                   # ---- peer_push_kernel
        for by in range(world - 1):
            q = (rank + 1 + by) % world
            q_len = min(per, nvec - q * per) if q * per < nvec else 0
            for bx in range(per // PUSH_VPC):
                for k in range(PUSH_U):
                    vi = bx * PUSH_VPC + tid + k * 256
                    vi = vi[vi < q_len]
                    stage[q][rank * per + vi] = bufs[rank][lo_vec + q * per + vi]
                    written[q][rank * per + vi] += 1
    assert all(w.max(initial=0) <= 1 for w in written)
    out = [b.copy() for b in bufs]
    for rank in range(world):                                   # ---- peer_reduce_kernel
        v_lo = rank * per
        length = min(per, nvec - v_lo) if v_lo < nvec else 0
        grid = max(1, (length + RED_VPC - 1) // RED_VPC)
        for bx in range(grid):
            for k in range(RED_U):
                vi = bx * RED_VPC + k * 256 + tid
                vi = vi[vi < length]
                acc = np.zeros(vi.shape)
                for r in range(world):
                    acc = acc + (bufs[rank][lo_vec + v_lo + vi] if r == rank else stage[rank][r * per + vi])
                for j in range(world):
                    out[(rank + j) % world][lo_vec + v_lo + vi] = acc / world
    return out


def _emulate_v1(bufs, lo_vec, nvec, stage_vecs, ctas):
    world = len(bufs)
    per = _per(nvec, world)
    chunks = per // 32
    stage = [np.full(stage_vecs, np.nan) for _ in range(world)]
    nwarps = ctas * 8
    lane = np.arange(32)
    for rank in range(world):                                   # ---- phase A
        units = chunks * (world - 1)
        for warp_g in range(nwarps):
            u0 = warp_g
            while u0 < units:
                for k in range(4):
                    u = u0 + k * nwarps
                    if u < units:
                        c, j = divmod(u, world - 1)
                        q = (rank + 1 + j) % world
                        vi = c * 32 + lane
                        gi = q * per + vi
                        ok = gi < nvec
                        stage[q][rank * per + vi[ok]] = bufs[rank][lo_vec + gi[ok]]
                u0 += 4 * nwarps
    out = [b.copy() for b in bufs]
    for rank in range(world):                                   # ---- phase C
        v_lo, v_hi = rank * per, min(nvec, rank * per + per)
        for warp_g in range(nwarps):
            c0 = warp_g
            while c0 * 32 < v_hi - v_lo:
                for k in range(2):
                    vi = (c0 + k * nwarps) * 32 + lane
                    vi = vi[v_lo + vi < v_hi]
                    acc = np.zeros(vi.shape)
                    for r in range(world):
                        acc = acc + (bufs[rank][lo_vec + v_lo + vi] if r == rank else stage[rank][r * per + vi])
                    for j in range(world):
                        out[(rank + j) % world][lo_vec + v_lo + vi] = acc / world
                c0 += 2 * nwarps
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("nvec", [1, 33, 2048, 2049, 3 * 2048 + 17, 40000])
def test_partition_arithmetic_reduces_every_vector_exactly_once(world, nvec):
    lib = _lib.load()
    stage_bytes = lib.ub200_peer_stage_bytes(C.c_int64(nvec * 8), C.c_int32(world))
    assert stage_bytes == _per(nvec, world) * world * 16
    rng = np.random.default_rng(world * 1000 + nvec)
    lo_vec, pad = 5, 7
    bufs = [rng.standard_normal(lo_vec + nvec + pad) for _ in range(world)]
    want = sum(b[lo_vec:lo_vec + nvec] for b in bufs) / world
    for name, outs in (("v2", _emulate_v2(bufs, lo_vec, nvec, stage_bytes // 16)),
                       ("v1", _emulate_v1(bufs, lo_vec, nvec, stage_bytes // 16, ctas=3))):
        for r, o in enumerate(outs):
            np.testing.assert_allclose(o[lo_vec:lo_vec + nvec], want, rtol=1e-12, atol=1e-12, err_msg=name)
            assert np.array_equal(o[:lo_vec], bufs[r][:lo_vec]) and np.array_equal(o[lo_vec + nvec:], bufs[r][lo_vec + nvec:])
            assert np.array_equal(o[lo_vec:lo_vec + nvec], outs[0][lo_vec:lo_vec + nvec])   # identical on all ranks


def test_argument_checks_do_not_need_a_gpu():
    lib = _lib.load()
    a = _lib.PeerAllreduceArgs()
    a.rank, a.world, a.offset, a.count = 0, 9, 0, 8
    assert lib.ub200_peer_allreduce(a, None) == -1                 # world > UB200_MAX_PEERS
    a.world, a.offset = 2, 4
    assert lib.ub200_peer_allreduce(a, None) == -1                 # offset not a multiple of 8 elements
    a.offset, a.count = 0, 0
    assert lib.ub200_peer_allreduce(a, None) == 0                  # empty slice: nothing to do
    assert lib.ub200_peer_flags_bytes() == 256
