"""All-reduce (SUM) of the config-4 gradient arena on N GPUs of one box: device time per call, max over ranks, for the
payloads the data-parallel train step can send -- whole arena fp32 / bf16, and the same bytes as 4 back-to-back buckets."""
import os, torch, torch.distributed as dist
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 33_554_432 + 8 * 64
def timed(fn, reps=20):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    buf = torch.zeros(N, dtype=dt, device=dev)
    es = buf.element_size()
    one = timed(lambda: dist.all_reduce(buf))
    q = N // 4
    four = timed(lambda: [dist.all_reduce(buf[i * q:(i + 1) * q]) for i in range(4)])
    small = timed(lambda: dist.all_reduce(buf[:q]))
    if rank == 0:
        bus = lambda nbytes, ms: nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
        print(f"{world} GPUs {name}: whole arena {N * es / 1e6:.0f} MB in {one:.3f} ms (busbw {bus(N * es, one):.0f} GB/s); "
              f"4 buckets back to back {four:.3f} ms; one bucket of {q * es / 1e6:.0f} MB {small:.3f} ms (busbw {bus(q * es, small):.0f} GB/s)", flush=True)
if rank == 0:
    f = torch.zeros(N, dtype=torch.float32, device=dev)
    def cast_rt():
        h = f.to(torch.bfloat16); f.copy_(h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cast_rt(); e0.record()
    for _ in range(10): cast_rt()
    e1.record(); torch.cuda.synchronize()
    print(f"fp32 -> bf16 -> fp32 round trip of the arena (the bf16 payload's extra passes): {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
