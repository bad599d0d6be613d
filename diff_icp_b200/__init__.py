"""diff_icp_b200 -- B200-native (sm_100a) implementation of the diffICP data-parallel hot path.

Layout mirrors the reference package (diffICP/{tools,core,api}) so that it drops in:

    tools/kernel.py      GaussKernel: the ten kernel reductions + coverage     (CUDA: csrc/ops_ksum.cuh)
    core/LDDMM.py        LDDMMModel: ODE / Shoot / trajloss / Optimize         (CUDA: csrc/ops_rhs.cuh, shooting.py)
    core/GMM.py          GaussianMixtureUnif: fused EM step                    (CUDA: csrc/em.cu)
    core/PSR.py          MultiPSR / DiffPSR outer alternation
    core/registrations.py, api/ICP_two_set.py, api/ICP_atlas.py

Host code is Python/PyTorch (device memory, streams, autograd plumbing, L-BFGS); all pairwise arithmetic runs in
hand-written CUDA kernels reached through the C ABI of libdicp_b200.so (include/dicp_b200.h).  There is no KeOps,
no Triton, no multi-backend dispatch and no CPU fallback.
"""

__version__ = "0.1.0"
