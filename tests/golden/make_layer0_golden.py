#!/usr/bin/env python
"""Writes tests/golden/layer0_activations.npz from the reference's own plaintext activations of encoder layer 0
(/root/reference/data/layer_0/**/allresults/*.csv, the values its encrypted run is meant to reproduce for the 5-token
test sentence, M/test/test_full_scheme.hpp:41-67,342) and data/selfoutput_linear.txt (BASELINE.json configs[0]).
The CSVs stay where they are; only the arrays the golden-gate tests read are stored (float64, compressed).
Run where /root/reference exists: python tests/golden/make_layer0_golden.py"""
import os

import numpy as np

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def csv(rel):
    return np.loadtxt(os.path.join(REF, rel), delimiter=",", dtype=np.float64)


L0 = "layer_0"
out = {
    # softmax stage: scores (5 tokens x [12 heads x 5 tokens]) -> probabilities            softmax.hpp:308-581
    "QKT": csv(L0 + "/Attention/BertSelfAttention/allresults/QKT.csv"),
    "aftsoftmax": csv(L0 + "/Attention/BertSelfAttention/allresults/aftsoftmax.csv"),
    # LayerNorm 1: residual sum -> normalised                                              layernorm.hpp:157-351
    "ln1_in": csv(L0 + "/Attention/SelfOutput/allresults/self_output_residual_connection_before_layernorm.csv"),
    "ln1_out": csv(L0 + "/Attention/SelfOutput/allresults/real_self_output.csv"),
    "ln1_gamma": csv(L0 + "/Attention/SelfOutput/parms/self_output_LayerNorm_weight.csv"),
    "ln1_beta": csv(L0 + "/Attention/SelfOutput/parms/self_output_LayerNorm_bias.csv"),
    # GELU: intermediate linear output -> activation                                       gelu_others.hpp:4-154
    "gelu_in": csv(L0 + "/Intermediate/allresults/intermediate_output_after_linear.csv"),
    "gelu_out": csv(L0 + "/Intermediate/allresults/real_intermediate_output.csv"),
    # LayerNorm 2                                                                          layernorm.hpp:353-547
    "ln2_in": csv(L0 + "/Output/allresults/final_output_residual_connection_before_layernorm.csv"),
    "ln2_out": csv(L0 + "/Output/allresults/real_final_output.csv"),
    "ln2_gamma": csv(L0 + "/Output/parms/final_output_LayerNorm_weight.csv"),
    "ln2_beta": csv(L0 + "/Output/parms/final_output_LayerNorm_bias.csv"),
    # config C1: the self-output linear layer's input (5 tokens x 768)                     test_ct_pt_matrix_mul.hpp
    "selfoutput_linear": np.loadtxt(os.path.join(REF, "selfoutput_linear.txt"), dtype=np.float64),
}
for k, v in out.items():
    print(k, v.shape, float(np.abs(v).max()))
np.savez_compressed(os.path.join(HERE, "layer0_activations.npz"), **out)
