"""BASELINE.json config 4: GPT2-XL beam search (beam 5) over 16 384 synthetic images, data-parallel over the GPUs of one box.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/run_config4.py [images] [new_tokens]
Every rank captions its contiguous range of image ids in micro-batches (Engine.caption_dataset: 256 // beam images per call on
the persistent decode kernel), then one NCCL all-gather of the caption tokens; time = CUDA events, max over ranks."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import clipcap_b200 as cc
from clipcap_b200 import synthetic, sharding

N_IMAGES = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cfg = cc.EngineConfig(max_images=64, max_beam=5, max_ctx=80)
eng = cc.Engine(cfg, local)
synthetic.load_synthetic(eng, 1234)
torch.cuda.empty_cache()
lo, hi = sharding.shard_range(N_IMAGES, rank, world)
# the shard's images, generated in blocks keyed by global image id (the same images whatever the number of GPUs)
blocks = []
for b0 in range(lo, hi, 64):
    n = min(64, hi - b0)
    blocks.append(synthetic.synthetic_images(64, cfg, seed=100000 + b0 // 64, device=dev)[:n])
images = torch.cat(blocks)
p = eng.gen_params("beam", T, stop_token=-1, max_stops=0, beam_size=5)


def run():
    return sharding.caption_images_sharded(eng, images, p, n_items=N_IMAGES)


warm = images[:128]
eng.caption_dataset(warm, p)            # graph capture, allocator
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tokens, lengths = run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    assert tokens.shape[0] == N_IMAGES
    print(json.dumps({"config": "GPT2-XL beam 5, %d images, %d new tokens" % (N_IMAGES, T), "n_gpus": world,
                      "seconds": float(ms) / 1e3, "captions_per_s": N_IMAGES / (float(ms) / 1e3),
                      "micro_batch_images": eng.micro_batch_for(p), "tokens_shape": list(tokens.shape),
                      "checksum": int(tokens.long().sum())}))
if world > 1:
    dist.destroy_process_group()
