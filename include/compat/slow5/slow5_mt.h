/* <slow5/slow5_mt.h>: the batch API twins live in slow5b200_file.h */
#include "slow5.h"
