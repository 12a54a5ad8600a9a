import sys, torch, faulthandler
faulthandler.dump_traceback_later(50, exit=True)
sys.path.insert(0, ".")
from tests import util
from uniter_b200 import ops, _lib
import uniter_b200.model as M

def sync(msg):
    torch.cuda.synchronize(); print("ok:", msg, flush=True)

for cfg, batch, name in ((util.TINY, util.tiny_batch(), "tiny"), (util.BASE_L1, util.c1_batch(False), "c1a")):
    print("=====", name, flush=True)
    model = util.make_model(cfg, util.make_state(cfg), torch.float16).eval()
    b = util.batch_to(batch, "cuda")
    meta = model._pack_meta(b["attn_masks"]); sync("meta T=%d" % meta["total"])
    model._weight_table(); sync("weight table")
    # monkeypatch lib calls to sync after each
    lib = M._bind()
    orig = {}
    for fn in ("ub200_embed_prep", "ub200_embed_gather_cast", "ub200_gemm", "ub200_embed_rows_fwd",
               "ub200_encoder_fwd", "ub200_gather_rows"):
        f = getattr(lib, fn)
        def wrap(*a, _f=f, _n=fn):
            rc = _f(*a); torch.cuda.synchronize(); print("   done", _n, rc, flush=True); return rc
        orig[fn] = f
        setattr(lib, fn, wrap)
    with torch.no_grad():
        out = model(b["input_ids"], b["position_ids"], b["img_feat"], b["img_pos_feat"], b["attn_masks"],
                    b["gather_index"], output_all_encoded_layers=False)
    sync("forward " + name)
    for fn, f in orig.items():
        setattr(lib, fn, f)
print("ALL DONE")
