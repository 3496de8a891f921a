"""B200-native (sm_100a) drop-in for the ALD MRI-reconstruction hot path of
10258392511/InverseProblemWithDiffusionModel.

The sub-packages mirror the reference's module paths for this path only:

    ncsn.linear_transforms            LinearTransform, i2k_complex, k2i_complex, generate_mask
    ncsn.linear_transforms.undersampling_fourier   RandomUndersamplingFourier, SENSE
    ncsn.linear_transforms.finite_diff             FiniteDiff
    ncsn.models                       get_sigmas, anneal_Langevin_dynamics
    ncsn.models.ncsnv2                NCSNv2, NCSNv2Deepest
    ncsn.models.proximal_op           L2Penalty, SingleCoil, Constrained, get_proximal
    ncsn.models.ALD_optimizers        ALDOptimizer, ALDUnconditionalSampler, ALDInvSegProximalRealImag, ALD2DTime
    sde.sampling                      AnnealedLangevinDynamics (the 'ald' corrector)
    chains                            chain sharding over GPUs + posterior mean / std reduction

Every compute call goes to hand-written CUDA in libipdm_b200.so through the C ABI of
include/ipdm_b200.h; there is no CPU or library fallback.
"""
__version__ = "0.1.0"
