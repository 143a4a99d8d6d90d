"""2-GPU micro-probe of the half-shard swap (prefix <-> suffix qubit) over NVLink: the three implementations of
qb_p2p_swapHalves, at several suffix positions.  torchrun --nproc-per-node 2 tools/p2p_probe.py [local_qubits]"""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quest_b200 import capi, quest_api as qa

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
capi.call("qb_bind_device", lr)
idbuf = (C.c_char * 128)()
if rank == 0:
    capi.call("qb_comm_get_unique_id", idbuf)
t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
dist.broadcast(t, src=0)
os.environ["QUEST_B200_NCCL_ID"] = bytes(t.cpu().tolist()).hex()
Q = qa.QuEST(qa.B200_LIB)
Q.initCustomQuESTEnv(1, 1, 0)
n = nl + world.bit_length() - 1
q = Q.createCustomQureg(n, 0, 1, 1, 0)
Q.initDebugState(q)
s = capi.qb_state()
s.amps, s.buffer, s.numAmpsPerNode, s.logNumAmpsPerNode = q.gpuAmps, q.gpuCommBuffer, 1 << nl, nl
s.rank, s.numQubits, s.logNumColsPerNode, s.isDensityMatrix = rank, n, 0, 0
lib = capi.lib()
pair = rank ^ 1
gib = 16 * (1 << nl) / 2 / 2**30
for mode in (0, 1, 2):
    lib.qb_p2p_set_swap_mode(mode)
    for sq in sorted({nl - 1, nl - 2, nl - 5, 16, 10, 3} & set(range(nl))):
        for _ in range(2):
            capi.check(lib.qb_p2p_swapHalves(C.byref(s), sq, pair), "swap")
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 4
        e0.record()
        for _ in range(reps):
            capi.check(lib.qb_p2p_swapHalves(C.byref(s), sq, pair), "swap")
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"mode {mode} suffix bit {sq:2d}: {ms.item():8.3f} ms  = {gib * 2**30 / ms.item() / 1e6:7.1f} GB/s per direction ({gib:.0f} GiB each way)", flush=True)
# an even number of swaps per (mode, bit) leaves the debug state intact: check through the API
amp = Q.getQuregAmp(q, 12345)
if rank == 0:
    print("amp[12345] after all swaps:", amp, "(debug state: 2*12345/10, (2*12345+1)/10)")
Q.destroyQureg(q); Q.finalizeQuESTEnv(); dist.barrier(); dist.destroy_process_group()
