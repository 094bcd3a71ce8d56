"""-m gpu, needs two GPUs (skipped otherwise): the C-ABI collective dpt_allgather_depth (include/dpt_b200.h) driven the way
a C/C++ host would drive it - with its own ncclComm_t, created here through ctypes on the NCCL that PyTorch ships - and
compared with the shards it was given."""
import ctypes as C
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _nccl():
    import nvidia.nccl  # the wheel PyTorch depends on

    path = glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*"))[0]
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib


def _worker(rank, world, uid_bytes, q):
    torch.cuda.set_device(rank)
    from muggled_dpt_b200 import _native as N

    nccl = _nccl()
    uid = _UniqueId()
    C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    b_local, H, W = 3, 20, 28
    g = torch.Generator().manual_seed(5)
    full = torch.randn(world * b_local, H, W, generator=g).to(torch.bfloat16)
    local = full[rank * b_local:(rank + 1) * b_local].cuda().contiguous()
    out = torch.empty_like(full, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = N.lib().dpt_allgather_depth(comm, C.c_void_p(local.data_ptr()), C.c_void_p(out.data_ptr()), local.numel(),
                                     N.DPT_BF16, stream)
    torch.cuda.synchronize()
    ok = rc == 0 and torch.equal(out.cpu(), full)
    nccl.ncclCommDestroy(comm)
    q.put((rank, bool(ok), int(rc)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allgather_depth_c_entry_two_ranks():
    import torch.multiprocessing as mp

    nccl = _nccl()
    uid = _UniqueId()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, C.string_at(C.byref(uid), 128), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, 0), (1, True, 0)], res


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_models_on_two_gpus_in_one_process():
    """one process, one handle per device: kernels opt into their shared-memory size per DEVICE (not per process), and
    every entry point runs on the handle's device whatever the caller's current device is"""
    import tempfile

    from muggled_dpt_b200 import make_dpt_from_state_dict
    from oracle import dpt_oracle as O

    sd = O.make_synthetic_state_dict("vits", seed=11)
    img = O.make_input(2, 252, 196, seed=9)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "depth_anything_v2_vits.pth")
        torch.save(sd, path)
        _, m0 = make_dpt_from_state_dict(path)
        _, m1 = make_dpt_from_state_dict(path)
    m0.to(device="cuda:0", dtype=torch.bfloat16)
    m1.to(device="cuda:1", dtype=torch.bfloat16)
    torch.cuda.set_device(0)  # current device stays 0 while the second model runs on device 1
    with torch.inference_mode():
        d1 = m1(img.to("cuda:1", torch.bfloat16))
        d0 = m0(img.to("cuda:0", torch.bfloat16))
        t1, g1 = m1.patch_embed(img.to("cuda:1", torch.bfloat16))
        taps1 = m1.imgencoder(t1, g1)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert d0.device.index == 0 and d1.device.index == 1 and taps1[0].device.index == 1
    assert torch.equal(d0.cpu(), d1.cpu())
