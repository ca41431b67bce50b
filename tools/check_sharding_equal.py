"""W-GPU output == 1-GPU output on the same image ids (pyramid item 4): captions of a fixed image set for
greedy, nucleus sampling (Philox streams keyed by global image id) and beam 5, sharded over however many ranks run this
script; rank 0 saves the gathered tokens.  Run it once plain and once under torchrun, then compare:

    python tools/check_sharding_equal.py gpurun_out/shard_n1.pt
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/check_sharding_equal.py gpurun_out/shard_n8.pt
    python tools/check_sharding_equal.py --compare gpurun_out/shard_n1.pt gpurun_out/shard_n8.pt [report.json]

Micro-batches are cut at the same global image ids whatever the number of ranks (every rank's range is a multiple of the
micro-batch), so the kernels see identical rows in identical launches and the tokens must be EQUAL, not merely close."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

if len(sys.argv) > 1 and sys.argv[1] == "--compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    rep = {"n_gpus": [a["world"], b["world"]]}
    ok = True
    for k in a["tokens"]:
        ta, tb = a["tokens"][k], b["tokens"][k]
        same = bool(torch.equal(ta, tb))
        rows = int((ta == tb).flatten(1).all(1).sum())
        rep[k] = {"images": ta.shape[0], "identical": same, "rows_identical": rows, "shape": list(ta.shape)}
        ok = ok and same
    rep["all_identical"] = ok
    print(json.dumps(rep, indent=1))
    if len(sys.argv) > 4:
        with open(sys.argv[4], "w") as f:
            json.dump(rep, f, indent=1)
    sys.exit(0 if ok else 1)

import torch.distributed as dist
import clipcap_b200 as cc
from clipcap_b200 import sharding, synthetic

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
T = 32
out = {}
for mode, n_images, mb, kw in (("greedy", 2048, 64, {}), ("sample", 2048, 64, dict(top_p=0.9, temperature=1.0, seed=3)),
                               ("beam", 1024, 32, dict(beam_size=5))):
    beam = kw.get("beam_size", 1)
    cfg = cc.EngineConfig(max_images=64, max_beam=beam, max_ctx=80)
    eng = cc.Engine(cfg, local)
    synthetic.load_synthetic(eng, 1234)
    lo, hi = sharding.shard_range(n_images, rank, world)
    assert (hi - lo) % mb == 0 and lo % mb == 0
    # images keyed by global block id: the same pixels whatever the number of ranks
    images = torch.cat([synthetic.synthetic_images(64, cfg, seed=100000 + b0 // 64, device=dev)[b0 % 64:b0 % 64 + mb]
                        for b0 in range(lo, hi, mb)])
    p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, **kw)
    tokens, lengths, _ = eng.caption_dataset(images, p, micro_batch=mb, first_row_id=lo)
    tokens, lengths = sharding.gather_captions(tokens, lengths, n_images)
    out[mode] = tokens.cpu()
    eng.close()
    del eng
    torch.cuda.empty_cache()
if rank == 0:
    torch.save({"world": world, "tokens": out}, sys.argv[1])
    print("saved", sys.argv[1], {k: tuple(v.shape) for k, v in out.items()})
if world > 1:
    dist.destroy_process_group()
