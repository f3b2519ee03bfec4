// align.cu -- placeholder until the GPU alignment stage lands (next milestone).
#include "align.h"

namespace opalb200 {
int align_database(DeviceDb*, const unsigned char*, int, unsigned char* const*, int, const int*, int, int, const int*, int,
                   OpalSearchResult*[], int) {
    set_error("OPAL_SEARCH_ALIGNMENT: GPU alignment stage not built yet");
    return OPAL_B200_ERR_CUDA;
}
}  // namespace opalb200
