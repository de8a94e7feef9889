"""yagre_mcmc_b200 -- B200-native batched-chain backend for the Metropolis-Hastings
hot path of rkutri/yagre-mcmc.  Same public chain / model / parameter / statistics
API as the reference (see the sub-packages), executed by hand-written sm_100a
kernels behind libyagre_b200.so (include/yagre_b200.h).  No CPU fallback."""
__version__ = "0.1.0"
