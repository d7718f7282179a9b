"""`from utils import inspect_net, load_from_caffe2_pkl` (reference utils.py:13-64, train.py:9,50)."""
from fastvideotagging_b200.utils import inspect_net, load_from_caffe2_pkl, caffe2_blobs_to_params      # noqa: F401
