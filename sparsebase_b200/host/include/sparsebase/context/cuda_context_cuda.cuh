// Same include path as the reference's src/sparsebase/context/cuda_context_cuda.cuh; the B200 host layer lives in sb200/.
#pragma once
#include "../../sb200/core.h"
