// rn_loss_inst_logits.cu -- one quarter of the flat loss kernel family (see rn_loss_kernel.cuh): rn_dispatch_loss_part<true, RnMatchI32>.
#include "rn_loss_kernel.cuh"

template void rn_dispatch_loss_part<true, RnMatchI32>(RN_LOSS_PART_ARGS);
