"""N=2 NCCL run of the sharded calibration / online range on real GPUs: results must equal the
single-GPU results on the whole batch bit for bit.  Needs two visible GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(device):
    from torch import nn
    from quantization.mxnet_b200 import model_zoo as Z
    from quantization.mxnet_b200.quantize import convert
    from quantization.mxnet_b200.quantize.initialize import qparams_init
    torch.manual_seed(7)
    net = Z.get_model("cifar_resnet20_v1", classes=10).eval().to(device)
    convert.convert_model(net, exclude=Z.default_exclusions(net, "cifar_resnet20_v1"))
    qparams_init(net)
    return net


def _data():
    g = torch.Generator().manual_seed(3)
    return [torch.randn(16, 3, 32, 32, generator=g) * (1 + 0.3 * i) for i in range(3)]


def _worker(rank, port, out):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
        torch.backends.cudnn.allow_tf32 = False
        from quantization.mxnet_b200 import dist as fqdist
        from quantization.mxnet_b200.quantize.distribution_calibrate import collect_feature_maps, kl_calibrate_all
        net = _build(dev)
        batches = _data()
        # --- KL calibration on shards
        net.disable_quantize()
        loader = [(fqdist.shard_batch(b), None) for b in batches]
        hist, fmax = collect_feature_maps(net, 2048, loader, dev)
        best = kl_calibrate_all(hist, 256, 256, 2048)
        blocks = net.collect_quantized_blocks()
        res = {"hist": np.stack([np.pad(hist[m], (0, 2049 - len(hist[m]))) for m in blocks]),
               "max": np.array([fmax[m] for m in blocks]), "best": best.cpu().numpy()}
        # --- online input quantisation with the global Kahan mean
        net.enable_quantize()
        net.quantize_input(True, online=True)
        fqdist.enable_data_parallel(net)
        with torch.no_grad():
            logits = net(fqdist.shard_batch(batches[0]).to(dev))
        res["cur"] = np.array([m.current_input_max.item() for m in blocks], np.float32)
        res["logits"] = logits.cpu().numpy()
        # --- offline inputs with range tracking: no collective in the forward, one all-gather in update_ema
        net.quantize_input(True, online=False)
        for m in blocks:
            m.input_max.data.fill_(1.5)
        with torch.no_grad():
            net(fqdist.shard_batch(batches[1]).to(dev))
        net.update_ema()
        res["ema"] = np.array([m.input_max.item() for m in blocks], np.float32)
        # --- the same step as a training step: update_ema() is deferred and the per-sample maxima ride in the tail
        #     of the gradient all-reduce (dist.GradBucket(net=...)): one collective, same input_max
        for m in blocks:
            m.input_max.data.fill_(1.5)
        params = [p for p in net.parameters() if p.requires_grad]
        bucket = fqdist.GradBucket(params, net=net)
        for step in range(2):           # the second step runs with the blocks' buffers inside the bucket's tail
            if step == 1:
                for m in blocks:
                    m.input_max.data.fill_(1.5)
            y = net(fqdist.shard_batch(batches[1]).to(dev))
            net.update_ema()
            assert bucket._deferred_ema == [0.9]
            y.sum().backward()
            bucket.all_reduce_mean()
            res["ema_bucket%d" % step] = np.array([m.input_max.item() for m in blocks], np.float32)
        res["grad0"] = params[0].grad.detach().cpu().numpy().copy()
        out.put((rank, res))
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_calibration_equals_single_gpu():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(v, dict) for v in results.values()), results

    # single-GPU reference on the whole batches
    from quantization.mxnet_b200.quantize.distribution_calibrate import collect_feature_maps, kl_calibrate_all
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    net = _build(dev)
    batches = _data()
    net.disable_quantize()
    hist, fmax = collect_feature_maps(net, 2048, [(b, None) for b in batches], dev)
    best = kl_calibrate_all(hist, 256, 256, 2048).cpu().numpy()
    blocks = net.collect_quantized_blocks()
    want_hist = np.stack([np.pad(hist[m], (0, 2049 - len(hist[m]))) for m in blocks])
    net.enable_quantize()
    net.quantize_input(True, online=True)
    with torch.no_grad():
        logits = net(batches[0].to(dev)).cpu().numpy()
    cur = np.array([m.current_input_max.item() for m in blocks], np.float32)
    net.quantize_input(True, online=False)
    for m in blocks:
        m.input_max.data.fill_(1.5)
    with torch.no_grad():
        net(batches[1].to(dev))
    net.update_ema()
    ema = np.array([m.input_max.item() for m in blocks], np.float32)
    for r in range(WORLD):
        assert results[r]["ema"][0] == ema[0]
        assert np.allclose(results[r]["ema"], ema, rtol=1e-5)
        # the first converted block sees identical inputs on both paths: bit-exact everything
        assert np.array_equal(results[r]["hist"][0], want_hist[0])
        assert results[r]["max"][0] == np.float32(fmax[blocks[0]])
        assert results[r]["cur"][0] == cur[0]
        # deeper layers: cuDNN may pick a different algorithm for a batch of 8 than for 16, so inputs can
        # differ by ulps; integer statistics must still agree to a handful of counts
        assert np.abs(results[r]["hist"] - want_hist).sum() <= 1e-4 * want_hist.sum()
        assert np.array_equal(results[r]["best"], best) or np.abs(results[r]["best"] - best).max() <= 2
        assert np.allclose(results[r]["cur"], cur, rtol=1e-5)
    # both ranks agree with each other exactly
    assert np.array_equal(results[0]["hist"], results[1]["hist"])
    assert np.array_equal(results[0]["best"], results[1]["best"])
    assert np.array_equal(results[0]["cur"], results[1]["cur"])
    assert np.array_equal(results[0]["ema"], results[1]["ema"])
    for r in range(WORLD):
        for step in range(2):
            assert results[r]["ema_bucket%d" % step][0] == results[r]["ema"][0]
            assert np.allclose(results[r]["ema_bucket%d" % step], results[r]["ema"], rtol=1e-6)
    assert np.array_equal(results[0]["ema_bucket1"], results[1]["ema_bucket1"])
    assert np.array_equal(results[0]["grad0"], results[1]["grad0"]) and np.abs(results[0]["grad0"]).sum() > 0
    got = np.concatenate([results[0]["logits"], results[1]["logits"]])
    assert np.abs(got - logits).max() <= 2e-2 * np.abs(logits).max()


# ---------------------------------------------------------------------------------------------------------------
# fake-BN batch statistics under data parallelism (convert_conv2d.py:148-153, convert.py:75-78)
# ---------------------------------------------------------------------------------------------------------------
def _build_fake_bn(device):
    from torch import nn
    from quantization.mxnet_b200 import model_zoo as Z
    from quantization.mxnet_b200.quantize import convert
    from quantization.mxnet_b200.quantize.initialize import qparams_init
    torch.manual_seed(7)
    net = Z.get_model("cifar_resnet20_v1", classes=10).eval().to(device)
    fn = {nn.Conv2d: convert.gen_conv2d_converter(weight_width=4, input_width=4, quant_type="channel", fake_bn=True),
          nn.Linear: convert.gen_dense_converter(weight_width=4, input_width=4, quant_type="channel"),
          nn.ReLU: None, nn.BatchNorm2d: convert.bypass_bn}
    convert.convert_model(net, exclude=Z.default_exclusions(net, "cifar_resnet20_v1"), convert_fn=fn)
    qparams_init(net)
    return net


def _fake_bn_state(net):
    blocks = [m for m in net.collect_quantized_blocks() if getattr(m, "running_mean", None) is not None]
    return {"mean": np.concatenate([m.running_mean.detach().cpu().numpy() for m in blocks]),
            "var": np.concatenate([m.running_var.detach().cpu().numpy() for m in blocks]),
            "cur_mean": np.concatenate([m.current_mean.cpu().numpy() for m in blocks]),
            "cur_var": np.concatenate([m.current_var.cpu().numpy() for m in blocks]),
            "first_c": blocks[0].out_channels}


def _fake_bn_worker(rank, port, out):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
        from quantization.mxnet_b200 import dist as fqdist
        net = _build_fake_bn(dev)
        fqdist.enable_data_parallel(net)
        net.quantize_input(True, online=True)
        res = {}
        for step, b in enumerate(_data()[:2]):
            with torch.no_grad():
                net(fqdist.shard_batch(b).to(dev))
            net.update_ema()
            res["step%d" % step] = _fake_bn_state(net)
        out.put((rank, res))
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_fake_bn_statistics_equal_the_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fake_bn_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(v, dict) for v in results.values()), results
    dev = torch.device("cuda", 0)
    net = _build_fake_bn(dev)
    net.quantize_input(True, online=True)
    for step, b in enumerate(_data()[:2]):
        with torch.no_grad():
            net(b.to(dev))
        net.update_ema()
        want = _fake_bn_state(net)
        c0 = want["first_c"]
        for key in ("mean", "var", "cur_mean", "cur_var"):
            # every rank holds the same running statistics, bit for bit
            assert np.array_equal(results[0]["step%d" % step][key], results[1]["step%d" % step][key]), key
            got = results[0]["step%d" % step][key]
            # ... and they are the single-GPU statistics of the global batch.  The first fake-BN block sees the
            # same input on both paths (only cuDNN's algorithm choice for N=8 vs N=16 can move its conv output by
            # ulps); deeper blocks inherit ulp-level input differences through re-quantisation
            assert np.allclose(got[:c0], want[key][:c0], rtol=2e-6, atol=1e-7), key
            assert np.allclose(got, want[key], rtol=5e-3, atol=1e-4), key
        # without the exchange the shard-local variance would be visibly off: the test has teeth
        assert np.abs(results[0]["step%d" % step]["cur_var"]).sum() > 0


# ---------------------------------------------------------------------------------------------------------------
# the C ABI's own collectives (include/fq.h fq_dist_*): a host that owns an ncclComm_t, no torch.distributed
# ---------------------------------------------------------------------------------------------------------------
def _nccl_library_path():
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
    return os.path.abspath(cands[0]) if cands else "libnccl.so.2"


def _c_abi_worker(rank, uid_queue, out):
    try:
        import ctypes
        import signal
        signal.alarm(240)           # a communicator that never forms must not outlive the test
        torch.cuda.set_device(rank)
        from quantization.mxnet_b200 import _ffi, ops
        from test_channel_stats_math import mean_close
        from oracle import build_c as C
        from oracle import fq_oracle as O
        path = _nccl_library_path()
        nccl = ctypes.CDLL(path)

        class UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]
        nccl.ncclGetUniqueId.argtypes = [ctypes.POINTER(UniqueId)]
        nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
        nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        uid = UniqueId()
        if rank == 0:
            assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
            for _ in range(WORLD - 1):
                uid_queue.put(ctypes.string_at(ctypes.byref(uid), 128))      # all 128 bytes (.internal stops at a NUL)
        else:
            ctypes.memmove(ctypes.byref(uid), uid_queue.get(timeout=120), 128)
        comm = ctypes.c_void_p()
        assert nccl.ncclCommInitRank(ctypes.byref(comm), WORLD, uid, rank) == 0
        lib = _ffi.load()
        _ffi.check_call(lib.fq_nccl_load(path.encode()))
        st = _ffi.current_stream()
        r = np.random.RandomState(17)
        # (1) online input range of the global batch: all-gather of per-sample maxima + Kahan mean
        x = np.abs(r.standard_normal((16, 8, 12, 12))).astype(np.float32)
        shard = torch.from_numpy(x[rank * 8:(rank + 1) * 8]).cuda()
        per = torch.empty(8, device="cuda")
        ops.input_range(shard, per_sample=per)
        per_all, cur = torch.empty(16, device="cuda"), torch.empty(1, device="cuda")
        a, b, c = _ffi.dl(per), _ffi.dl(per_all), _ffi.dl(cur)
        _ffi.check_call(lib.fq_dist_input_range(a.ptr, b.ptr, c.ptr, comm, st))
        want_cur, want_per = O.input_range(x)
        assert np.array_equal(per_all.cpu().numpy(), want_per) and np.float32(cur.item()) == want_cur
        # (2) histogram counts of 2 batches: integer sum over ranks + float32 fold in batch order (32-bit wire format)
        mx = torch.tensor([float(x.max())], device="cuda")
        counts = torch.zeros(2, 2049, dtype=torch.int32, device="cuda")
        for bi in range(2):
            ops.hist_nonzero(shard * (1.0 - 0.3 * bi), mx, 2048, counts[bi], promotion="nep50")
        hist = torch.zeros(2049, device="cuda")
        a, b = _ffi.dl(counts), _ffi.dl(hist)
        _ffi.check_call(lib.fq_dist_hist_fold(a.ptr, b.ptr, 1, None, comm, st))
        want = sum(O.histogram_counts(x * np.float32(1.0 - 0.3 * bi), 2048, x.max(), "nep50").astype(np.float32)
                   for bi in range(2))
        assert np.array_equal(hist.cpu().numpy()[:2048], want[:2048]) and int(counts.abs().sum()) == 0
        # (3) fake-BN statistics of the global batch from the ranks' records
        y = (r.standard_normal((16, 6, 9, 9)) * 2 + 1).astype(np.float32)
        rec = torch.empty(6, 4, dtype=torch.float64, device="cuda")
        ops.channel_stats(torch.from_numpy(y[rank * 8:(rank + 1) * 8]).cuda(), parts=rec, finish=False)
        rec_all = torch.empty(WORLD, 6, 4, dtype=torch.float64, device="cuda")
        mean, var = torch.empty(6, device="cuda"), torch.empty(6, device="cuda")
        a, b, c, d = _ffi.dl(rec), _ffi.dl(rec_all), _ffi.dl(mean), _ffi.dl(var)
        _ffi.check_call(lib.fq_dist_channel_stats(a.ptr, b.ptr, c.ptr, d.ptr, comm, st))
        wm, wv = C.channel_stats(y)
        assert mean_close(mean.cpu().numpy(), wm, y)
        assert np.abs(var.cpu().numpy().view(np.int32).astype(np.int64) - wv.view(np.int32).astype(np.int64)).max() <= 2
        # (4) plain reductions: MAX of first-batch ranges, SUM of a gradient bucket
        t = torch.tensor([1.0 + rank, 5.0 - rank], device="cuda")
        a = _ffi.dl(t)
        _ffi.check_call(lib.fq_dist_all_reduce(a.ptr, 1, comm, st))
        assert t.tolist() == [2.0, 5.0]
        gbuf = torch.full((1000,), float(rank + 1), device="cuda")
        a = _ffi.dl(gbuf)
        _ffi.check_call(lib.fq_dist_all_reduce(a.ptr, 0, comm, st))
        assert bool((gbuf == 3.0).all())
        torch.cuda.synchronize()
        nccl.ncclCommDestroy(comm)
        out.put((rank, "ok"))
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c_abi_collectives_with_a_raw_nccl_communicator():
    ctx = mp.get_context("spawn")
    out, uid_queue = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_c_abi_worker, args=(r, uid_queue, out), daemon=True) for r in range(WORLD)]
    for p in procs:
        p.start()
    try:
        results = sorted(out.get(timeout=300) for _ in procs)
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    assert results == [(0, "ok"), (1, "ok")], results
