#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* c, cudaGraphConditionalHandle h, int maxit) {
    int v = ++(*c);
    cudaGraphSetConditional(h, v < maxit ? 1u : 0u);
}
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
    printf("addnode %d\n", (int)e);
    cudaGraph_t bg = p.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(st, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    body<<<1, 1, 0, st>>>(d, h, 5);
    e = cudaStreamEndCapture(st, nullptr);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", (int)e);
    cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
    int hv; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost); printf("count %d err %d\n", hv, (int)cudaGetLastError());
}
