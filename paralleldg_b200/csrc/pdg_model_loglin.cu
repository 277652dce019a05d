// loglin instantiations of the MH, logl[0] and batch-scorer kernels (see pdg_model.inl)
#define PDG_TU_MODEL PDG_MODEL_LOGLIN
#define PDG_TU_NAME loglin
#define PDG_TU_IS_LOGLIN 1
#include "pdg_model.inl"
