"""Data-parallel plumbing around the training step (train_lidog.py:227-231, 288: SyncBN conversion + Lightning's
`strategy='ddp'`, i.e. `DistributedDataParallel` with its defaults).

`DistributedDataParallel(broadcast_buffers=True)` re-broadcasts rank 0's module buffers before every forward.  For
MinkUNet34BEV that is ~190 tensors (running mean / variance / num_batches_tracked of 62 + 2 batch norms): torch
flattens them, broadcasts, and copies every tensor back with its own kernel -- measured 1.7 ms per 42 ms step at 2
GPUs (profiles/r02_f_ddp_ab_2gpu.txt), all of it before the first kernel of the forward.  `FlatBuffers` keeps the
semantics (rank 0's buffers are authoritative at the start of every forward) and drops the cost: the buffers are
re-bound once as views of one flat tensor per dtype, and a step broadcasts those two tensors."""
import torch
import torch.distributed as dist


class FlatBuffers:
    def __init__(self, module):
        groups = {}
        for mod in module.modules():
            for name, buf in list(mod._buffers.items()):
                if buf is not None:
                    groups.setdefault((buf.dtype, buf.device), []).append((mod, name, buf))
        self.flat, self.views = [], []
        for (dtype, device), items in groups.items():
            align = max(1, 16 // torch.empty((), dtype=dtype).element_size())  # slices stay 16-byte aligned
            offsets, total = [], 0
            for _, _, b in items:
                offsets.append(total)
                total += -(-b.numel() // align) * align
            flat = torch.zeros(total, dtype=dtype, device=device)
            for (mod, name, b), o in zip(items, offsets):
                view = flat[o:o + b.numel()].view(b.shape)
                view.copy_(b)
                mod._buffers[name] = view
                self.views.append((mod, name, view.data_ptr()))
            self.flat.append(flat)

    def intact(self):
        """False once something (`module.to`, `.float()`, ...) re-bound a buffer away from the flat storage."""
        return all(mod._buffers[name] is not None and mod._buffers[name].data_ptr() == ptr for mod, name, ptr in self.views)

    def broadcast(self, src=0, group=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return
        for f in self.flat:
            dist.broadcast(f, src, group=group)


def wrap(net, device_ids=None, buffers="flat", **ddp_kwargs):
    """-> (ddp_module, FlatBuffers or None).  buffers: 'flat' (above), 'ddp' (torch's own per-tensor broadcast),
    'off' (no broadcast: training-mode results are identical, the non-zero ranks' running statistics of the
    un-synchronised 2D-head batch norms then drift from rank 0's)."""
    if buffers not in ("flat", "ddp", "off"):
        raise ValueError(f"buffers must be 'flat', 'ddp' or 'off', not {buffers!r}")
    fb = FlatBuffers(net) if buffers == "flat" else None
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=device_ids, broadcast_buffers=(buffers == "ddp"),
                                                    **ddp_kwargs)
    return ddp, fb
