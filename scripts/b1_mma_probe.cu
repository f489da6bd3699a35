#include <cstdint>
__global__ void k(const uint32_t* a, const uint32_t* b, int* c) {
    uint32_t A[4], B[2]; int C[4] = {0,0,0,0};
    for (int i=0;i<4;++i) A[i]=a[threadIdx.x*4+i];
    for (int i=0;i<2;++i) B[i]=b[threadIdx.x*2+i];
    asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]) : "r"(A[0]),"r"(A[1]),"r"(A[2]),"r"(A[3]),"r"(B[0]),"r"(B[1]));
    for (int i=0;i<4;++i) c[threadIdx.x*4+i]=C[i];
}
