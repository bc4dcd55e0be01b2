// ogb_internal.h -- shared by the host and device halves of libogb.so.
#ifndef OGB_INTERNAL_H_
#define OGB_INTERNAL_H_

#include <cstdarg>
#include <cstdint>
#include <cstdlib>

#include "../../include/ogb.h"

void ogb_set_error(const char *fmt, ...);

#endif
