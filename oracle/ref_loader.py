"""Import the UNMODIFIED reference (CoderWZW/ARLib) from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  /root/reference exists in
the builder container and NOT on the GPU box, so nothing that runs under
``-m gpu``, ``smoke()`` or ``bench.py`` may call into this module; it is used by
``oracle/make_golden.py`` (to freeze golden vectors) and by the CPU tests that
cross-check ``oracle/port.py`` against the live reference when it is mounted.

Shims, all on the harness side, reference files untouched:
  1. ``.cuda()`` -> identity (the reference hard-codes ``.cuda()``;
     recommender/LightGCN.py:31,38-39,43,210,215) so it runs on CPU.
  2. ``sys.argv`` reset before the argparse parsers (conf/recommend_parser.py:4-34).
  3. ``torch.sparse.FloatTensor`` deprecation warnings silenced.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("ARLIB_REFERENCE_ROOT", "/root/reference")

# module names the reference owns at top level; they would collide with
# nothing in this repo (ours live under arlib_b200.*), but are purged on exit
_REF_TOP = ("util", "recommender", "conf", "attack", "ARLib")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "recommender", "LightGCN.py"))


@contextlib.contextmanager
def reference_modules():
    """Context manager: puts /root/reference first on sys.path, applies the
    shims, yields a namespace with the imported reference modules, then
    restores everything."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    import torch

    saved_path = list(sys.path)
    saved_argv = list(sys.argv)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in _REF_TOP}
    for k in list(saved_mods):
        del sys.modules[k]
    saved_tcuda = torch.Tensor.cuda
    saved_mcuda = torch.nn.Module.cuda
    sys.path.insert(0, REF_ROOT)
    sys.argv = [saved_argv[0] if saved_argv else "oracle"]
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = types.SimpleNamespace()
            import util.DataLoader as _dl
            import util.sampler as _sampler
            import util.loss as _loss
            import util.algorithm as _alg
            import util.metrics as _metrics
            import util.tool as _tool
            import conf.recommend_parser as _rp
            import recommender.LightGCN as _lg
            import recommender.NGCF as _ngcf
            import recommender.SimGCL as _sim
            import recommender.XSimGCL as _xsim
            ns.DataLoader = _dl.DataLoader
            ns.sampler = _sampler
            ns.loss = _loss
            ns.algorithm = _alg
            ns.metrics = _metrics
            ns.tool = _tool
            ns.recommend_parse_args = _rp.recommend_parse_args
            ns.LightGCN = _lg
            ns.NGCF = _ngcf
            ns.SimGCL = _sim
            ns.XSimGCL = _xsim
            yield ns
    finally:
        torch.Tensor.cuda = saved_tcuda
        torch.nn.Module.cuda = saved_mcuda
        sys.argv = saved_argv
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOP]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def make_args(ns, **overrides):
    """The reference's argparse namespace with defaults, then overrides."""
    args = ns.recommend_parse_args()
    args.load = False
    args.save = False
    for k, v in overrides.items():
        setattr(args, k, v)
    return args
