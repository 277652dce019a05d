// uniform instantiations of the MH, logl[0] and batch-scorer kernels (see pdg_model.inl)
#define PDG_TU_MODEL PDG_MODEL_UNIFORM
#define PDG_TU_NAME uniform
#include "pdg_model.inl"
