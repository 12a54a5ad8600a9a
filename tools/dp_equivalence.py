"""N-rank NCCL data-parallel step == 1-rank step with N x gradient accumulation over the same shards
(the reference's own "gradient accumulation emulates multi-gpu" equivalence, README.md:115).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dp_equivalence.py [--graph] [--peer]

--peer: the slices travel through the library's own NVLink peer-memory kernel (PeerExchange,
csrc/peer.cu) instead of NCCL; with --graph the exchange kernels are nodes of the step's one graph.

Every rank r computes the gradients of shard r (dropout off) and the GradientReducer averages them
over NCCL (overlapped chunked all-reduce of arena slices, exactly the bench path; with --graph the
whole step incl. the all-reduces is a captured CUDA graph).  Rank 0 then recomputes, alone, the
gradients of EVERY shard with accumulation into the arena and divides by N.  The two must agree to
within the 16-bit rounding of the arena (the N-rank path rounds each shard's gradient to 16 bit before
averaging, the accumulated path rounds the running sum).  Prints one JSON line on rank 0; exit code
0 iff the check passed."""
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
    use_graph = "--graph" in sys.argv
    use_peer = "--peer" in sys.argv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from uniter_b200 import distributed as ubd
    from uniter_b200.arena import GradArena
    from uniter_b200.graphed import GraphedStep
    from uniter_b200.heads import UniterForMLM
    from uniter_b200.model import UniterConfig, register_lengths
    from uniter_b200.synth import pad_mlm_index, synth_batch

    torch.manual_seed(0)
    cfg = UniterConfig(2000, hidden_size=256, num_hidden_layers=4, num_attention_heads=4,
                       intermediate_size=1024, max_position_embeddings=64)
    model = UniterForMLM(cfg, 128).to(dev, torch.bfloat16).eval()          # dropout off
    ubd.broadcast_parameters(model, root=0)
    arena = GradArena.attach(model)

    def shard(r):
        b = pad_mlm_index(synth_batch(16, 6, 14, 4, 20, seed=50 + r, img_dim=128, vocab_size=2000, mlm_prob=0.3), 16)
        lens = [a + c for a, c in zip(b["txt_lens"], b["num_bbs"])]
        return {k: v for k, v in b.items() if torch.is_tensor(v)}, lens

    def loss_fn(b):
        return (model(b).sum() * b["mlm_inv_n"]).squeeze()

    reducer = ubd.GradientReducer(model, overlap_chunks=2, transport="peer" if use_peer else "nccl")
    if use_peer:
        reducer.peer.timeout_ms = 5000
    hb, lens = shard(rank)
    if use_graph:
        step = GraphedStep(model, loss_fn, token_bucket=64, reducer=reducer,
                           reducer_mode="in-graph" if use_peer else "split")
        step(hb, lens)
        step(hb, lens)                                # a second replay: same result, nothing accumulates
    else:
        b = {k: v.to(dev) for k, v in hb.items()}
        register_lengths(b["attn_masks"], lens, prefix=True)
        model.zero_grad(set_to_none=True)
        reducer.backward_and_reduce(loss_fn(b))
    torch.cuda.synchronize()
    peer_err = reducer.peer.error_word() if use_peer else 0
    got = arena.flat.float().clone()
    # every rank must hold the same reduced gradients
    ref0 = got.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(ref0, got))
    flags = torch.tensor([1.0 if same else 0.0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    ok, rec = True, {}
    if rank == 0:
        model.zero_grad(set_to_none=True)
        arena.flat.zero_()
        acc = torch.zeros_like(got)
        for r in range(world):                       # 1 rank, N x accumulation over the same shards
            hb_r, lens_r = shard(r)
            b = {k: v.to(dev) for k, v in hb_r.items()}
            register_lengths(b["attn_masks"], lens_r, prefix=True)
            model.zero_grad(set_to_none=True)
            loss_fn(b).backward()
            acc += arena.flat.float()
        want = acc / world
        num = (got - want).norm().item()
        den = want.norm().item()
        rec = {"world": world, "graph": use_graph, "rel_err": num / den, "grad_norm": den,
               "max_abs": (got - want).abs().max().item(), "ranks_identical": bool(flags.item() == 1.0),
               "arena_elements": int(arena.numel), "transport": "peer" if use_peer else "nccl",
               "peer_error_word": peer_err}
        ok = rec["rel_err"] < 5e-3 and rec["ranks_identical"] and peer_err == 0
        print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
