"""Correctness + bandwidth check of the NVLink peer-memory all-reduce (csrc/peer.cu, PeerExchange)
against NCCL, on N ranks of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29561 tools/peer_check.py [--mb 220]

1. random 16-bit buffers, slices of awkward sizes (8 elements ... the whole buffer): the result must
   equal the fp32 mean of all ranks' inputs rounded once to 16 bit (computed with an NCCL fp32
   all-reduce), on every rank, bit for bit across ranks; repeated calls; bf16 and fp16;
2. the same call captured in a CUDA graph and replayed (the epochs live in device memory);
3. timing of the whole buffer (default 220 MB = the UNITER-base arena) for several CTA counts, next to
   ncclAllReduce(AVG) of the same buffer: algorithmic GB/s = bytes / time.
Rank 0 prints one JSON line; exit code 0 iff every comparison passed and no flag wait expired."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    mb = 220
    if "--mb" in sys.argv:
        mb = int(sys.argv[sys.argv.index("--mb") + 1])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from uniter_b200 import distributed as ubd

    n = mb * (1 << 20) // 2 // 8 * 8
    ok, rec = True, {"world": world, "elements": n}
    for dtype in (torch.bfloat16, torch.float16):
        flat = torch.zeros(n, device=dev, dtype=dtype)
        px = ubd.PeerExchange(flat, timeout_ms=5000)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        worst = 0.0
        for lo, hi in [(0, 8), (8, 272), (1024, 1024 + 8 * 4097), (0, 1 << 20), (40, n), (0, n)]:
            hi = min(hi, n)
            for rep in range(3):
                src = torch.randn(hi - lo, device=dev, generator=g) * (1.0 + rank)
                flat[lo:hi] = src.to(dtype)
                before = flat.clone()
                want = flat[lo:hi].float()
                dist.all_reduce(want, op=dist.ReduceOp.SUM)
                want = (want * (1.0 / world)).to(dtype)
                dist.barrier()
                px.all_reduce(lo, hi, max_ctas=(-1, 0, 7)[rep])    # the three forms: copy engines / short-lived CTAs / persistent
                torch.cuda.synchronize()
                got = flat[lo:hi]
                # two ranks: a + b is order independent -> bit exact; more ranks: NCCL's fp32 summation
                # order differs from the kernel's fixed 0..world-1 order by at most the last fp32 bit
                exact = bool(torch.equal(got, want)) if world == 2 else \
                    bool(torch.allclose(got.float(), want.float(), rtol=8e-3, atol=1e-6))
                ident = got.clone()
                dist.broadcast(ident, src=0)
                exact = exact and bool(torch.equal(ident, got))       # every rank holds the same bits
                untouched = bool(torch.equal(flat[:lo], before[:lo]) and torch.equal(flat[hi:], before[hi:]))
                d = (got.float() - want.float()).abs().max().item()
                worst = max(worst, d)
                if not (exact and untouched):
                    ok = False
                    rec.setdefault("failures", []).append([str(dtype), lo, hi, rep, d, untouched])
        # graph replay
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        src = torch.randn(n, device=dev, generator=g)
        with torch.cuda.stream(s):
            px.all_reduce(0, n)               # warm (eager)
            torch.cuda.synchronize()
            with torch.cuda.graph(gr, stream=s):
                px.all_reduce(0, n)
        torch.cuda.current_stream().wait_stream(s)
        for rep in range(3):
            flat.copy_((src * (rep + 1)).to(dtype))
            want = flat.float()
            dist.all_reduce(want, op=dist.ReduceOp.SUM)
            want = (want * (1.0 / world)).to(dtype)
            gr.replay()
            torch.cuda.synchronize()
            if not (torch.equal(flat, want) if world == 2 else
                    torch.allclose(flat.float(), want.float(), rtol=8e-3, atol=1e-6)):
                ok = False
                rec.setdefault("failures", []).append([str(dtype), "graph", rep])
        rec["max_abs_diff_%s" % str(dtype).split(".")[-1]] = worst
        err = px.error_word()
        if err:
            ok = False
            rec["peer_error_word"] = err
        if dtype == torch.bfloat16:
            # ---- timing
            times = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for ctas in (-1, 0, 148):        # copy engines / short-lived CTAs / persistent
                for _ in range(3):
                    px.all_reduce(0, n, max_ctas=ctas)
                dist.barrier()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    px.all_reduce(0, n, max_ctas=ctas)
                e1.record()
                torch.cuda.synchronize()
                times["peer_form_%d" % ctas] = e0.elapsed_time(e1) / 10
            for _ in range(3):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            e1.record()
            torch.cuda.synchronize()
            times["nccl_avg"] = e0.elapsed_time(e1) / 10
            t = torch.tensor([times[k] for k in sorted(times)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rec["ms"] = {k: round(v, 4) for k, v in zip(sorted(times), t.tolist())}
            rec["algbw_GBps"] = {k: round(n * 2 / (v * 1e-3) / 1e9, 1) for k, v in rec["ms"].items()}
            err = px.error_word()
            if err:
                ok = False
                rec["peer_error_word"] = err
        px.close()
        del px, flat
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item() == 1.0)
    rec["ok"] = ok
    if rank == 0:
        print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
