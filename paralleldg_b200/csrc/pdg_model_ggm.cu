// ggm instantiations of the MH, logl[0] and batch-scorer kernels (see pdg_model.inl)
#define PDG_TU_MODEL PDG_MODEL_GGM
#define PDG_TU_NAME ggm
#define PDG_TU_IS_GGM 1
#include "pdg_model.inl"
