// step_resident (resident.cuh): all calls between two snapshots in one cooperative launch.  See launch.h.
#include "launch.h"

namespace fdtd {

template <typename T>
cudaError_t resident_blocks_per_sm(int* per_sm) {
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, step_resident<T>, RES_THREADS, 0);
}

template <typename T>
cudaError_t launch_step_resident(const StepParams<T>& P, T* h1, T* ex1, T* ey1, const ResSources<T>& R, long long first, int ns,
                                 int blocks, cudaStream_t st) {
    StepParams<T> p = P;
    ResSources<T> r = R;
    void* args[] = {&p, &h1, &ex1, &ey1, &r, &first, &ns};
    return cudaLaunchCooperativeKernel((void*)step_resident<T>, dim3((unsigned)blocks), dim3(RES_THREADS), args, 0, st);
}

template cudaError_t resident_blocks_per_sm<double>(int*);
template cudaError_t resident_blocks_per_sm<float>(int*);
template cudaError_t launch_step_resident<double>(const StepParams<double>&, double*, double*, double*, const ResSources<double>&,
                                                  long long, int, int, cudaStream_t);
template cudaError_t launch_step_resident<float>(const StepParams<float>&, float*, float*, float*, const ResSources<float>&,
                                                 long long, int, int, cudaStream_t);

}  // namespace fdtd
