/* Stand-in for <jemalloc/jemalloc.h>: the reference only includes it (src/main.c:3)
 * and then calls plain malloc/free, so glibc's allocator is a drop-in. Test infrastructure only. */
#pragma once
#include <stdlib.h>
