// style_kspace.h — as style_pair.h, for KSpaceStyle(key,Class) (pppm_intel.h:18-22, pppm_disp_intel.h:18-22 of the reference)
#include "pppm_intel.h"
#include "pppm_disp_intel.h"
