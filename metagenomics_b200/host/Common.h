// Common.h -- shared definitions of the host-side drop-in classes (metagenomics_b200/host).
//
// Mirrors the role of MetaGenomics/Common.h:31-37 (integer typedefs of identical width, so that the
// class signatures below match the reference headers) without its tuning macros, which do not touch
// the overlap-graph build. Errors of the GPU library surface as OgbFailure instead of the
// reference's print-and-exit(0) (Common.h:47).
#ifndef OGB_HOST_COMMON_H_
#define OGB_HOST_COMMON_H_

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ogb.h"

using namespace std;

typedef unsigned char UINT8;
typedef unsigned short UINT16;
typedef short INT16;
typedef unsigned long UINT32;		// 64-bit on LP64, as in the reference (Common.h:34)
typedef long INT32;
typedef unsigned long long UINT64;
typedef long long INT64;

class OgbFailure : public std::runtime_error
{
public:
	int code;
	OgbFailure(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

// Turns a non-zero libogb status into an exception carrying ogb_last_error().
inline void ogbCheck(int status, const char *where)
{
	if (status != OGB_OK)
		throw OgbFailure(status, std::string(where) + ": " + ogb_last_error());
}

#endif
