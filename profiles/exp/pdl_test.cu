// Does griddepcontrol.wait wait for the COMPLETION of the primary grid when the primary triggers early?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void primary(int* data, int n, int early_trigger) {
  if (early_trigger) asm volatile("griddepcontrol.launch_dependents;");
  long long t0 = clock64();
  while (clock64() - t0 < 40000) {}  // ~20 us
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) data[i] = 1;
}
__global__ void secondary(const int* data, int n, int* missing) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ((volatile const int*)data)[i] != 1) atomicAdd(missing, 1);
}
int main() {
  int n = 1 << 20, *data, *missing;
  cudaMalloc(&data, n * 4); cudaMalloc(&missing, 4);
  cudaStream_t s; cudaStreamCreate(&s);
  for (int early = 0; early < 2; ++early)
    for (int pdl = 0; pdl < 2; ++pdl) {
      int total = 0;
      for (int rep = 0; rep < 20; ++rep) {
        cudaMemsetAsync(data, 0, n * 4, s); cudaMemsetAsync(missing, 0, 4, s);
        primary<<<n / 256, 256, 0, s>>>(data, n, early);
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(n / 256); cfg.blockDim = dim3(256); cfg.stream = s;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = pdl;
        cfg.attrs = a; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, secondary, (const int*)data, n, missing);
        int m; cudaMemcpyAsync(&m, missing, 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s);
        total += m;
      }
      printf("early_trigger=%d pdl_attr=%d missing(sum over 20 reps)=%d  err=%s\n", early, pdl, total, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
