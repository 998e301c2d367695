#define FXK_SL_K 4
#include "k_sl.inc"
