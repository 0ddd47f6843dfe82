// Stands in for M/source/ckks_evaluator.h on the facade include path.  The reference's softmax.hpp and
// single_att_block.hpp include that header but use nothing from it (its CKKSEvaluator class belongs to
// an unused code path that also needs the client-side Encryptor), so nothing is declared here.
#pragma once
#include "seal/seal.h"
