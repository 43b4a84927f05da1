"""How fast do many small volatile stores into a PEER's memory go?  (the record-mailing pattern of te_finish_step)
    torchrun --nproc-per-node 2 scratch/probe_remote_stores.py
Rank 0 times kernels that store into rank 1's buffer (and into its own, for reference)."""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genjax_b200.inference.pf_dist import SymmArena
from genjax_b200.runtime import cabi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
arena = SymmArena(1 << 22, dev, dist.group.WORLD)
core = cabi.core()
fn = core.gjb_remote_store_probe
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
fn.restype = C.c_int
st = cabi.stream_ptr(dev)
dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    def t(ptr, ctas, dests, words, coal, reps=50):
        for _ in range(5):
            fn(ptr, ctas, dests, words, coal, 7 << 32, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(reps):
            fn(ptr, ctas, dests, words, coal, (i + 8) << 32, st)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps
    peer, own = arena.ptrs[1 % world], arena.ptrs[0]
    print("us per launch (512 CTAs; stores = ctas x dests x words)")
    for dests in (1, 2, 4, 8):
        print(f"  dests={dests} words=3: peer {t(peer, 512, dests, 3, 0):7.2f}   own {t(own, 512, dests, 3, 0):7.2f}   "
              f"peer, one CTA set coalesced {t(peer, 512, dests, 3, 1):7.2f}", flush=True)
    print(f"  empty-ish launch (dests=0): {t(peer, 512, 0, 3, 0):7.2f}")
dist.barrier(); torch.cuda.synchronize()
dist.destroy_process_group()
