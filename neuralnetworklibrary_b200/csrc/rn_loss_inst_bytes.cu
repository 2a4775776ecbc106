// rn_loss_inst_bytes.cu -- one quarter of the flat loss kernel family (see rn_loss_kernel.cuh): rn_dispatch_loss_part<false, RnMatchU8NC>.
// Only part of builds with -DRN_EXPERIMENTAL (the byte-map chain of rn_loss_step).
#include "rn_loss_kernel.cuh"

template void rn_dispatch_loss_part<false, RnMatchU8NC>(RN_LOSS_PART_ARGS);
