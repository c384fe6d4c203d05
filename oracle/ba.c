/* ba.c -- bundle-adjustment oracle (placeholder until the BA row is built). TEST INFRASTRUCTURE ONLY. */
#include "oracle.h"
