__global__ void k_exp(const double* x, double* y) { int i = blockIdx.x * blockDim.x + threadIdx.x; y[i] = exp(x[i]); }
__global__ void k_div(const double* x, double* y) { int i = blockIdx.x * blockDim.x + threadIdx.x; y[i] = x[i] / x[i + 1]; }
__global__ void k_sincos(const double* x, double* y) { int i = blockIdx.x * blockDim.x + threadIdx.x; double s, c; sincos(x[i], &s, &c); y[i] = s * c; }
__global__ void k_sqrt(const double* x, double* y) { int i = blockIdx.x * blockDim.x + threadIdx.x; y[i] = sqrt(x[i]); }
