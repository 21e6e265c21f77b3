/* <slow5/slow5_press.h>: the codec interface twins live in slow5b200_press.h */
#include "slow5.h"
