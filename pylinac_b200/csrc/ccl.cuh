// Union-find on a parent array in global memory (shared by locate.cu and edges.cu): roots are the smallest index of a component
// (= skimage / scipy label order), hooking by atomicMin, path halving.
#pragma once

namespace epid {

__device__ __forceinline__ int gl_find(int* parent, int i) {
    while (true) {
        const int p = parent[i];
        if (p == i) return i;
        const int gp = parent[p];
        if (gp != p) atomicMin(&parent[i], gp);      // path halving; parents only ever decrease, so a concurrent hook is never lost
        i = p;
    }
}

__device__ __forceinline__ void gl_union(int* parent, int a, int b) {
    while (true) {
        a = gl_find(parent, a);
        b = gl_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // hook the larger root under the smaller one
        const int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

}  // namespace epid
