"""`mx.model.load_checkpoint` / `save_checkpoint` (validation.py:22, train.py:83 via do_checkpoint).

`prefix-%04d.params` is the MXNet NDArray-dict file (`arg:` / `aux:` keys; fastvideotagging_b200.params_io).
`prefix-symbol.json` holds the arguments `create_r3d` was called with — the graph is fully determined by them (net.py:110-170)
— instead of MXNet's operator-level graph JSON, which only an MXNet can execute."""
import json

from fastvideotagging_b200 import params_io
from fastvideotagging_b200.net import create_r3d

from . import ndarray as nd


def _sym_to_json(sym):
    return {"fvt_b200_symbol": "create_r3d", "num_class": sym.num_class, "model_depth": sym.model_depth,
            "final_temporal_kernel": sym.pool[0], "final_spatial_kernel": sym.pool[1], "bn_mom": sym.bn_mom}


def save_checkpoint(prefix, epoch, symbol, arg_params, aux_params):
    if symbol is not None:
        with open("%s-symbol.json" % prefix, "w") as fh:
            json.dump(_sym_to_json(symbol), fh)
    conv = lambda d: {k: (v.asnumpy() if hasattr(v, "asnumpy") else v) for k, v in d.items()}     # noqa: E731
    return params_io.save_checkpoint(prefix, epoch, conv(arg_params), conv(aux_params))


def load_checkpoint(prefix, epoch):
    with open("%s-symbol.json" % prefix) as fh:
        js = json.load(fh)
    if js.get("fvt_b200_symbol") != "create_r3d":
        raise NotImplementedError("%s-symbol.json is an MXNet operator graph; rebuild the symbol with net.create_r3d(...) and load "
                                  "the .params file with mx.nd.load" % prefix)
    sym = create_r3d(js["num_class"], no_bias=True, model_depth=js["model_depth"], final_spatial_kernel=js["final_spatial_kernel"],
                     final_temporal_kernel=js["final_temporal_kernel"], bn_mom=js["bn_mom"])
    arg, aux = params_io.load_checkpoint(prefix, epoch)
    return sym, {k: nd.array(v) for k, v in arg.items()}, {k: nd.array(v) for k, v in aux.items()}
