"""mizuroute_b200 -- B200-native reach-routing solver behind the C ABI of include/mizuroute_b200.h.

    capi       ctypes binding of libmizuroute_b200.so (built in-tree by build.py with nvcc for sm_100a)
    route      Router: host mirror of the reference's main_route interface (step, batch, async batch, state, remap)
    multi      DomainSet: tributary / mainstem domains across GPUs, NCCL hand-off
    partition  the reference's tributary/mainstem decomposition
    network    RiverNetwork / RouteParams / RouteOptions containers
    synth      seeded synthetic river networks and runoff (BASELINE.json configurations)
    casefiles  writes stand-alone cases in the reference's file formats (control file, namelist, NetCDF-3)

There is no CPU fallback: without the CUDA library or a CUDA device every entry point fails loudly.
"""
