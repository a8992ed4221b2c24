// Internal interface between watershed.cu (pre-processing, generic structuring elements, C ABI) and
// watershed_fast.cu (the persistent 6-connected engine).
#pragma once
#include <stdint.h>

int64_t b2v_wsf_workspace_bytes(int64_t nz, int64_t ny, int64_t nx);
// stages: 1 INIT | 2 COST_CONVERGE | 4 LABEL_BEGIN | 8 LABEL_CONVERGE | 16 FINISH (watershed_fast.cu)
int b2v_wsf_run(int stages, const uint16_t* img, const int16_t* markers, int64_t nz, int64_t ny, int64_t nx, int mode,
                int frozen_lo, int frozen_hi, int16_t* labels, uint8_t* ambiguous, int with_set, void* workspace,
                void* stream, int* rounds_io, int allow_key32);
