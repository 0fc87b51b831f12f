import os, sys, importlib, time
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist, numpy as np, fir_b200
synth = importlib.import_module("fast-image-recognition_b200.synth")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N, Q, k = 100000, 10000, 10
g, gl, q, ql = synth.make_split(N, Q, 512, 1000, "l2", seed=rank)
gd, qd = torch.from_numpy(g).to(dev), torch.from_numpy(q).to(dev)
fir_b200.normalize_rows(gd); fir_b200.normalize_rows(qd)
gal = fir_b200.Gallery(gd, torch.from_numpy(gl).to(dev), "l2", index_offset=rank*N, stream=torch.cuda.current_stream().cuda_stream)
gdd = torch.empty((world*Q, k), dtype=torch.float32, device=dev); gii = torch.empty((world*Q, k), dtype=torch.int32, device=dev)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else 'barrier'
for it in range(10):
    if mode == 'barrier': dist.barrier(); torch.cuda.synchronize()
    if mode in ('flush', 'flush_nobar'): flush.fill_(1)
    if mode == 'flush': dist.barrier()
    t0 = time.perf_counter()
    e0 = ev(); idx, dd = gal.search(qd, k=k); e1 = ev()
    dist.all_gather_into_tensor(gdd, dd); dist.all_gather_into_tensor(gii, idx); e2 = ev()
    mi, md = fir_b200.merge_topk(gdd.view(world, Q, k), gii.view(world, Q, k), stream=torch.cuda.current_stream().cuda_stream); e3 = ev()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if rank == 0 and it >= 3:
        print("search %.3f  gather %.3f  merge %.3f  total %.3f ms | host enqueue %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e0.elapsed_time(e3), 1e3*(t1-t0)))
dist.destroy_process_group()
