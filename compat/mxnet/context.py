"""`mx.gpu(i)` / `mx.cpu()` (reference train_simple_r3d.py:33, train.py:54, validation.py:23)."""
import torch


class Context:
    def __init__(self, device_type, device_id=0):
        self.device_type, self.device_id = device_type, int(device_id)

    @property
    def device(self):
        return torch.device("cuda", self.device_id) if self.device_type == "gpu" else torch.device("cpu")

    def __repr__(self):
        return "%s(%d)" % (self.device_type, self.device_id)

    def __eq__(self, other):
        return isinstance(other, Context) and (self.device_type, self.device_id) == (other.device_type, other.device_id)

    def __hash__(self):
        return hash((self.device_type, self.device_id))


def gpu(device_id=0):
    return Context("gpu", device_id)


def cpu(device_id=0):
    return Context("cpu", device_id)


def current_context():
    return cpu()


def one_device(ctx):
    """The single CUDA device behind a context / list of contexts.  This framework runs one process per GPU
    (torch.distributed + NCCL for the gradient sum), so several contexts in one process are refused loudly."""
    lst = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]
    if len(lst) != 1:
        raise NotImplementedError(
            "%d contexts in one process: this framework runs ONE process per GPU.  Launch with `python -m torch.distributed.run "
            "--nproc-per-node N <script> --gpus <LOCAL_RANK>` — gradients are summed over the processes by NCCL inside "
            "Trainer.step / Module.fit, where the reference's kvstore sums them over contexts." % len(lst))
    if lst[0].device_type != "gpu":
        raise RuntimeError("the R(2+1)D hot path runs on sm_100a only: pass --gpus <id> (no CPU fallback)")
    return lst[0].device
