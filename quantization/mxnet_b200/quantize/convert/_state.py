"""Packed per-net state (B200 execution detail, not part of the reference's API).

The reference keeps one tiny NDArray per block for every range and statistic it tracks (``input_max``, ``act_max``,
``running_mean``, ``running_var`` and the ``current_*`` values beside them; convert_conv2d.py:113-141,
convert.py:66-78) and updates them with one op each.  Here every kind lives in ONE pair of arenas (state vector,
current vector) and the per-block tensors are views into them, so that

  * ``net.update_ema()`` is one launch per kind instead of one (or two) per block,
  * data-parallel exchanges are one collective for the whole net,
  * device pointers never change after conversion: a CUDA graph captured at any time keeps reading and writing the
    tensors ``update_ema`` sees.

The arenas are built eagerly by ``convert_model`` and rebuilt right after ``net.to()/.cuda()/...`` (which replaces
the parameters' storage).  A rebuild that would be needed while a captured graph is alive raises instead of
silently re-pointing the tensors the graph holds.
"""
import types

import torch

# (state parameter, current buffer, "current plays a host scalar" -- convert.py:70 vs :75-78)
KINDS = (("input_max", "current_input_max", True), ("act_max", "current_act_max", True),
         ("running_mean", "current_mean", False), ("running_var", "current_var", False))


def _owners(blocks, state_name):
    return [m for m in blocks if getattr(m, state_name, None) is not None]


def _intact(arena, owners, state_name, current_name):
    if arena is None or arena["ids"] != tuple(map(id, owners)):
        return False
    state, cur = arena["state"], arena["current"]
    if state.device != getattr(owners[0], state_name).device:
        return False
    s0, c0 = state.data_ptr(), cur.data_ptr()
    for m, off in zip(owners, arena["offsets"]):
        c = getattr(m, current_name, None)
        if c is None or getattr(m, state_name).data.data_ptr() != s0 + 4 * off or c.data_ptr() != c0 + 4 * off:
            return False
    return True


def pack(net, blocks=None):
    """Make every block's state a view of the net's arenas; returns {state_name: arena dict}.  Cheap when nothing
    moved (pointer comparisons)."""
    blocks = net.collect_quantized_blocks() if blocks is None else blocks
    arenas = net.__dict__.setdefault("_fq_arenas", {})
    for state_name, current_name, _ in KINDS:
        owners = _owners(blocks, state_name)
        if not owners:
            arenas.pop(state_name, None)
            continue
        if _intact(arenas.get(state_name), owners, state_name, current_name):
            continue
        if net.__dict__.get("_fq_graph_live", False):
            raise RuntimeError("the packed %s state of this net would have to move, but a captured CUDA graph still "
                               "points at it: convert the net and move it to its device BEFORE capturing" % state_name)
        dev = getattr(owners[0], state_name).device
        sizes = [getattr(m, state_name).numel() for m in owners]
        offsets = [0]
        for n in sizes[:-1]:
            offsets.append(offsets[-1] + n)
        with torch.no_grad():
            state = torch.cat([getattr(m, state_name).data.reshape(-1).to(dev, torch.float32) for m in owners])
            cur_parts = []
            for m, n in zip(owners, sizes):
                c = getattr(m, current_name, None)
                cur_parts.append(torch.zeros(n, dtype=torch.float32, device=dev) if c is None or c.numel() != n
                                 else c.detach().reshape(-1).to(dev, torch.float32))
            cur = torch.cat(cur_parts)
        for m, off, n in zip(owners, offsets, sizes):
            p = getattr(m, state_name)
            p.data = state[off:off + n].view(p.shape)
            view = cur[off:off + n].view(p.shape)
            if current_name in m._buffers:
                m._buffers[current_name] = view
            else:
                m.__dict__.pop(current_name, None)
                m.register_buffer(current_name, view, persistent=False)
        arenas[state_name] = {"state": state, "current": cur, "ids": tuple(map(id, owners)), "offsets": offsets,
                              "sizes": sizes}
    _pack_stats_records(net, blocks, arenas)
    return arenas


def _pack_stats_records(net, blocks, arenas):
    """float64 [sum Cout, 4] arena of the fake-BN {n, S1, S2, K} records (csrc/fq_stats.cu), one slice per block."""
    owners = _owners(blocks, "running_mean")
    if not owners:
        arenas.pop("stats_parts", None)
        return
    dev = owners[0].running_mean.device
    rec = arenas.get("stats_parts")
    total = sum(m.running_mean.numel() for m in owners)
    if rec is not None and rec["ids"] == tuple(map(id, owners)) and rec["parts"].device == dev and \
            rec["parts"].shape[0] == total:
        return
    parts = torch.zeros(total, 4, dtype=torch.float64, device=dev)
    off = 0
    for m in owners:
        n = m.running_mean.numel()
        m.__dict__["_fq_stats_parts"] = parts[off:off + n]
        off += n
    arenas["stats_parts"] = {"parts": parts, "ids": tuple(map(id, owners)), "gathered": None}


def _apply_and_repack(self, fn, *args, **kwargs):
    """net._apply (what .to()/.cuda()/.float() call): the parameters get new storage, so the views are rebuilt."""
    out = type(self)._apply(self, fn, *args, **kwargs)
    if "_fq_arenas" in self.__dict__:
        pack(self)
    return out


def install(net):
    """Called by convert_model: pack now and after every move of the net."""
    pack(net)
    net.__dict__["_apply"] = types.MethodType(_apply_and_repack, net)
