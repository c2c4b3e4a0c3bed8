// Same include path as the reference's src/sparsebase/format/cuda_csc_cuda.cuh; the B200 host layer lives in sb200/.
#pragma once
#include "../../sb200/format.h"
