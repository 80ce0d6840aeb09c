// kdf_host.cpp — host-side (CPU) parts of libkdf_sm100.so: BGZF/BAM decode into
// packed pinned batches.  Filled in by the host-reader milestone.
#include "../../include/kdf.h"
