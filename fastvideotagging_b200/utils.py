"""Mirror of the importable, hot-path-adjacent part of reference utils.py: `load_from_caffe2_pkl` (utils.py:13-55,
called at train.py:50).  The OpenCV helpers of that file (`test_clip`, :65-90) are data plumbing, out of scope."""
from .params_io import load_from_caffe2_pkl, caffe2_blobs_to_params  # noqa: F401


def inspect_net(net):
    """Reference utils.py:58-64."""
    print("name %s" % getattr(net, "name", "r3d"))
    print("===========%d of arg============" % len(net.list_arguments()))
    print(net.list_arguments())
    print("===========%d of aux============" % len(net.list_auxiliary_states()))
    print(net.list_auxiliary_states())
