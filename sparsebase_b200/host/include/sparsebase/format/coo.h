// Same include path as the reference's src/sparsebase/format/coo.h; the B200 host layer lives in sb200/.
#pragma once
#include "../../sb200/format.h"
