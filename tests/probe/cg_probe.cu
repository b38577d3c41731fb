// Probe: CUDA graph conditional WHILE nodes (nested) on this driver, and the per-iteration latency of a loop body.
#include <cuda_runtime.h>
#include <cstdio>
struct St { int inner, outer, total; };
__global__ void k_inner(St* s, cudaGraphConditionalHandle h, int limit){ s->total++; int c = ++s->inner; cudaGraphSetConditional(h, c < limit ? 1u : 0u); }
__global__ void k_outer_begin(St* s, cudaGraphConditionalHandle hin){ s->inner = 0; cudaGraphSetConditional(hin, 1u); }
__global__ void k_outer_end(St* s, cudaGraphConditionalHandle hout, int limit){ int c = ++s->outer; cudaGraphSetConditional(hout, c < limit ? 1u : 0u); }
#define CK(x) do{cudaError_t e=(x); if(e){printf("ERR %s -> %s\n", #x, cudaGetErrorString(e)); return 1;}}while(0)
int main(){
  cudaStream_t s; CK(cudaStreamCreate(&s));
  St* d; CK(cudaMalloc(&d, sizeof(St))); CK(cudaMemset(d, 0, sizeof(St)));
  cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle hout; CK(cudaGraphConditionalHandleCreate(&hout, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams po = { cudaGraphNodeTypeConditional };
  po.conditional.handle = hout; po.conditional.type = cudaGraphCondTypeWhile; po.conditional.size = 1;
  cudaGraphNode_t nout; CK(cudaGraphAddNode(&nout, g, nullptr, 0, &po));
  cudaGraph_t gout = po.conditional.phGraph_out[0];
  // outer body: begin kernel -> inner while -> end kernel
  cudaGraphConditionalHandle hin; CK(cudaGraphConditionalHandleCreate(&hin, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNode_t nb;
  { cudaKernelNodeParams kp = {}; void* args[] = {&d, &hin}; kp.func = (void*)k_outer_begin; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    CK(cudaGraphAddKernelNode(&nb, gout, nullptr, 0, &kp)); }
  cudaGraphNodeParams pi = { cudaGraphNodeTypeConditional };
  pi.conditional.handle = hin; pi.conditional.type = cudaGraphCondTypeWhile; pi.conditional.size = 1;
  cudaGraphNode_t nin; CK(cudaGraphAddNode(&nin, gout, &nb, 1, &pi));
  cudaGraph_t gin = pi.conditional.phGraph_out[0];
  int inner_limit = 5, outer_limit = 1000;
  { cudaKernelNodeParams kp = {}; void* args[] = {&d, &hin, &inner_limit}; kp.func = (void*)k_inner; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    cudaGraphNode_t n; CK(cudaGraphAddKernelNode(&n, gin, nullptr, 0, &kp)); }
  { cudaKernelNodeParams kp = {}; void* args[] = {&d, &hout, &outer_limit}; kp.func = (void*)k_outer_end; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    cudaGraphNode_t n; CK(cudaGraphAddKernelNode(&n, gout, &nin, 1, &kp)); }
  cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CK(cudaGraphLaunch(ex, s)); CK(cudaStreamSynchronize(s));
  CK(cudaMemset(d, 0, sizeof(St)));
  cudaEventRecord(e0, s); CK(cudaGraphLaunch(ex, s)); cudaEventRecord(e1, s); CK(cudaStreamSynchronize(s));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  St h; CK(cudaMemcpy(&h, d, sizeof(St), cudaMemcpyDeviceToHost));
  printf("outer %d (expect 1000) total inner %d (expect 5000) time %.3f ms -> %.2f us per kernel node\n", h.outer, h.total, ms, 1e3*ms/(h.total + 2*h.outer));
  return 0;
}
