#define FXK_SL_K 2
#include "k_sl.inc"
