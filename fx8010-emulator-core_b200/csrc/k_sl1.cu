#define FXK_SL_K 1
#include "k_sl.inc"
