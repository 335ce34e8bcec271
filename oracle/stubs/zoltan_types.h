/* oracle/stubs: stand-in for Zoltan's id types (test infrastructure only) */
#pragma once
typedef unsigned int ZOLTAN_ID_TYPE;
typedef ZOLTAN_ID_TYPE *ZOLTAN_ID_PTR;
#define ZOLTAN_OK 0
#define ZOLTAN_WARN 1
#define ZOLTAN_FATAL (-1)
