// Same include path as the reference's src/sparsebase/converter/converter_order_two.h; the B200 host layer lives in sb200/.
#pragma once
#include "../../sb200/format.h"
