// The U-Net handle (shared by the forward plan, unet.cu, and the backward plan, unet_bwd.cu).
#pragma once
#include "net.cuh"

struct b200ad_unet : b200ad::NetBase {
  b200ad_unet_config cfg;
  struct Backward* bwd = nullptr;   // built lazily by b200ad_unet_backward (unet_bwd.cu)
};

namespace b200ad {
void release_backward(b200ad_unet* h);   // frees h->bwd (defined next to the Backward type, unet_bwd.cu)
}
