// core/utils/ThreadManager.h -- present only so that reference scripts that include it keep compiling.
// The kT/dT worker-thread hand-shake of the reference (src/core/utils/ThreadManager.h) does not exist here:
// the rebuild and the force/integration steps run in order on one CUDA stream.
#pragma once
