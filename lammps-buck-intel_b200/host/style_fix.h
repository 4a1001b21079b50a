// style_fix.h — as style_pair.h, for FixStyle(key,Class)
#include "fix_nve_intel.h"
