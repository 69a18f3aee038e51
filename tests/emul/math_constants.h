/* shadows CUDA's <math_constants.h> when kernels are compiled with g++ (cuda_on_cpu.h defines what they use) */
