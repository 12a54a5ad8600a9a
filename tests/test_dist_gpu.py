"""GPU, single-rank NCCL group: the data-parallel gradient exchange on real CUDA tensors
(chunked backward + overlapped arena all-reduce + the small-parameter bucket), SURVEY.md §8e.
With one rank the mean all-reduce is the identity, so the reduced gradients must equal those of a
plain backward — what is exercised is the plumbing the multi-GPU bench relies on."""
import os
import socket

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gradient_reducer_chunked_backward_single_rank_nccl():
    import torch.distributed as dist
    from uniter_b200 import distributed as ubd
    from uniter_b200.heads import UniterForMLM
    from uniter_b200.synth import synth_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        model = UniterForMLM(util.tiny_config(), 64).to("cuda", torch.bfloat16).eval()   # dropout off
        ubd.broadcast_parameters(model, root=0)
        batch = util.batch_to(synth_batch(6, 5, 12, 3, 9, seed=4, img_dim=64, vocab_size=2000, mlm_prob=0.3), "cuda")
        model(batch).mean().backward()
        ref = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        assert len(ref) > 40
        model.zero_grad(set_to_none=True)
        red = ubd.GradientReducer(model, overlap_chunks=2)
        red.backward_and_reduce(model(batch).mean())
        torch.cuda.synchronize()
        for n, p in model.named_parameters():
            if n in ref:
                assert p.grad is not None, n
                assert torch.allclose(p.grad.float(), ref[n].float(), atol=2e-3, rtol=2e-2), n
        # the encoder-layer gradients were reduced in place inside the arena
        q = model.uniter.encoder.layer[0].attention.self.query.weight
        assert q.grad.data_ptr() == model.uniter._ensure_arena()[0].view(q).data_ptr()
    finally:
        dist.destroy_process_group()
