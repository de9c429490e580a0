// Geometry of a stride-1 gather-GEMM convolution on NHWC tensors (shared by the SIMT and tcgen05 paths).
#pragma once
struct ConvGeom {
    int B, Hi, Wi, Ci;   // tensor gathered from
    int Ho, Wo, Co;      // tensor produced
    int KH, KW, ph, pw;  // window; input row = ho + kh - ph
    int ldx, ldy;        // pixel-row pitch (floats) of input / output
    int K;               // KH*KW*Ci
    long M;              // B*Ho*Wo
};
