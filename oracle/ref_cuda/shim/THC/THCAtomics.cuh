// TEST INFRASTRUCTURE ONLY - stand-in for <THC/THCAtomics.cuh>: float atomicAdd is a CUDA builtin.
#pragma once
