"""arlib_b200 -- B200-native (sm_100a) graph-CF train + full-rank evaluation hot path
of CoderWZW/ARLib, behind the reference's recommender / util interface.

  arlib_b200.recommender.{LightGCN,NGCF,SimGCL,XSimGCL}   drop-in recommender classes
  arlib_b200.util.{DataLoader,sampler,loss,algorithm,metrics,FileIO}   reference util surface
  arlib_b200.ops / graph / encoder / engine / evaluator   the device path (ctypes -> libagcf.so)

There is no CPU fallback: every compute entry point raises if libagcf.so or a CUDA
device is missing.
"""
__version__ = "0.1.0"
