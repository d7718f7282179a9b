"""`mx.callback.do_checkpoint` / `Speedometer` (train.py:83-84)."""
import logging
import time


def do_checkpoint(prefix, period=1):
    """epoch_end_callback: writes `prefix-symbol.json` and `prefix-%04d.params` (arg:/aux: NDArray dict)."""
    period = int(max(1, period))

    def _callback(iter_no, sym, arg, aux):
        if (iter_no + 1) % period == 0:
            from . import model
            model.save_checkpoint(prefix, iter_no + 1, sym, arg, aux)
    return _callback


class Speedometer:
    def __init__(self, batch_size, frequent=50, auto_reset=True):
        self.batch_size, self.frequent = batch_size, frequent
        self.tic, self.last = None, 0

    def __call__(self, param):
        count = param.nbatch
        if self.tic is None or count < self.last:
            self.tic, self.last = time.time(), count
            return
        if count % self.frequent == 0 and count != self.last:
            speed = self.frequent * self.batch_size / max(time.time() - self.tic, 1e-9)
            msg = "Epoch[%d] Batch [%d]\tSpeed: %.2f samples/sec" % (param.epoch, count, speed)
            if param.eval_metric is not None:
                msg += "\t%s=%f" % param.eval_metric.get()
            logging.info(msg)
            self.tic, self.last = time.time(), count
