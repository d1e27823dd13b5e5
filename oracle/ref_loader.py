"""Import the UNMODIFIED reference (CoderWZW/ARLib).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is looked up at
``$ARLIB_REFERENCE_ROOT``, then ``/root/reference`` (builder container), then the
staged byte-for-byte copy ``oracle/_ref/`` (``oracle/make_ref.py``; git-ignored,
shipped to the GPU box like the built library).  Users: ``oracle/make_golden*.py``
(freeze golden vectors), the CPU tests that cross-check ``oracle/port.py`` against
the live reference, the drop-in tests that run the reference's OWN driver / attack
code against ``arlib_b200.recommender.*`` and ``bench.py``'s baselines.  Nothing
under ``arlib_b200/`` imports this module.

Shims, all on the harness side, reference files untouched:
  1. ``cuda=False``: ``.cuda()`` -> identity (the reference hard-codes ``.cuda()``;
     recommender/LightGCN.py:31,38-39,43,210,215) so it runs on CPU.  With
     ``cuda=True`` the shim is NOT installed: the reference runs its own eager
     PyTorch + cuSPARSE path on the GPU.
  2. ``sys.argv`` reset before the argparse parsers (conf/recommend_parser.py:4-34).
  3. ``torch.sparse.FloatTensor`` deprecation warnings silenced.
  4. ``random.sample(set, k)`` (util/tool.py:84-92 and every fakeUserInject) raises
     on Python >= 3.11 -> sample from ``sorted(set)`` (only when ``callers=True``).
  5. ``dropin=True``: ``recommender.{LightGCN,NGCF,SimGCL,XSimGCL}`` resolve to the
     arlib_b200 classes (the re-export shim of INTEGRATION.md section 2, done in
     ``sys.modules`` instead of editing files); everything else -- ARLib.py, attack/*,
     util/* -- is the reference's own code.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, "_ref")


def _resolve_root():
    cands = [os.environ.get("ARLIB_REFERENCE_ROOT"), "/root/reference", STAGED]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "recommender", "LightGCN.py")):
            return c
    return "/root/reference"


REF_ROOT = _resolve_root()

# module names the reference owns at top level; they would collide with
# nothing in this repo (ours live under arlib_b200.*), but are purged on exit
_REF_TOP = ("util", "recommender", "conf", "attack", "ARLib")
_DROPIN = ("LightGCN", "NGCF", "SimGCL", "XSimGCL")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "recommender", "LightGCN.py"))


def is_staged_copy() -> bool:
    return os.path.abspath(REF_ROOT) == os.path.abspath(STAGED)


@contextlib.contextmanager
def reference_modules(cuda=False, dropin=False, callers=False):
    """Context manager: puts the reference first on sys.path, applies the shims,
    yields a namespace with the imported reference modules, then restores
    everything.  ``callers``: also import ARLib, conf.attack_parser and make
    ``ns.attack(kind, name)`` available."""
    if not available():
        raise RuntimeError("reference not found at %s (run python -m oracle.make_ref in the builder container)" % REF_ROOT)
    import random
    import torch

    saved_path = list(sys.path)
    saved_argv = list(sys.argv)
    saved_mods = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in _REF_TOP}
    for k in list(saved_mods):
        del sys.modules[k]
    saved_tcuda = torch.Tensor.cuda
    saved_mcuda = torch.nn.Module.cuda
    saved_sample = random.sample
    sys.path.insert(0, REF_ROOT)
    sys.argv = [saved_argv[0] if saved_argv else "oracle"]
    if not cuda:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if callers:
        def _sample(population, k, **kw):
            if isinstance(population, (set, frozenset)):
                population = sorted(population)
            return saved_sample(population, k, **kw)
        random.sample = _sample
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = types.SimpleNamespace()
            if dropin:
                import importlib
                import recommender as _pkg          # the reference's (namespace) package
                for name in _DROPIN:
                    mod = importlib.import_module("arlib_b200.recommender." + name)
                    sys.modules["recommender." + name] = mod
                    setattr(_pkg, name, mod)
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
            if callers:
                import importlib
                import ARLib as _arlib
                import conf.attack_parser as _ap
                ns.ARLib = _arlib.ARLib
                ns.attack_parse_args = _ap.attack_parse_args
                ns.attack = lambda kind, name: getattr(importlib.import_module("attack.%s.%s" % (kind, name)), name)
            yield ns
    finally:
        torch.Tensor.cuda = saved_tcuda
        torch.nn.Module.cuda = saved_mcuda
        random.sample = saved_sample
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


def make_attack_args(ns, **overrides):
    args = ns.attack_parse_args()
    for k, v in overrides.items():
        setattr(args, k, v)
    return args
