// The dense scoring kernel for fp32 operands on the TF32 tensor path (pack mode 3, alad_mrsw_fwd_args.operand_format = 1):
// the same source as mrsw_fwd.cu with kind::tf32 MMAs, compiled as its own translation unit (see the note there).
#define ALAD_MMA_TF32 1
#include "mrsw_fwd.cu"
