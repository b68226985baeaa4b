// placeholder, replaced below
#include "orc_core.h"
namespace orc { LightMap* lightmap_create(const slb_lightmap_desc*, int, int, int, int, int) { return nullptr; } }
