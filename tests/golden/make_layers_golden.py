#!/usr/bin/env python
"""Writes tests/golden/layers_1_11_activations.npz: the reference's own plaintext activations around the LayerNorm and
GELU stages of encoder layers 1..11 (/root/reference/data/layer_k/**/allresults/*.csv + the LayerNorm parameters), the
same files tests/golden/make_layer0_golden.py reads for layer 0.  Stored as float32 (2 MB instead of 4.3 MB; the gates
that read them have tolerances >= 1e-3).  Run where /root/reference exists: python tests/golden/make_layers_golden.py"""
import os

import numpy as np

REF = "/root/reference/data"
HERE = os.path.dirname(os.path.abspath(__file__))


def csv(rel):
    return np.loadtxt(os.path.join(REF, rel), delimiter=",", dtype=np.float64).astype(np.float32)


out = {}
for k in range(1, 12):
    L = "layer_%d" % k
    p = "l%d_" % k
    out[p + "ln1_in"] = csv(L + "/Attention/SelfOutput/allresults/self_output_residual_connection_before_layernorm.csv")
    out[p + "ln1_out"] = csv(L + "/Attention/SelfOutput/allresults/real_self_output.csv")
    out[p + "ln1_gamma"] = csv(L + "/Attention/SelfOutput/parms/self_output_LayerNorm_weight.csv")
    out[p + "ln1_beta"] = csv(L + "/Attention/SelfOutput/parms/self_output_LayerNorm_bias.csv")
    out[p + "gelu_in"] = csv(L + "/Intermediate/allresults/intermediate_output_after_linear.csv")
    out[p + "gelu_out"] = csv(L + "/Intermediate/allresults/real_intermediate_output.csv")
    out[p + "ln2_in"] = csv(L + "/Output/allresults/final_output_residual_connection_before_layernorm.csv")
    out[p + "ln2_out"] = csv(L + "/Output/allresults/real_final_output.csv")
    out[p + "ln2_gamma"] = csv(L + "/Output/parms/final_output_LayerNorm_weight.csv")
    out[p + "ln2_beta"] = csv(L + "/Output/parms/final_output_LayerNorm_bias.csv")
    print(L, {n[len(p):]: (v.shape, float(np.abs(v).max())) for n, v in out.items() if n.startswith(p) and ("_in" in n)})
np.savez_compressed(os.path.join(HERE, "layers_1_11_activations.npz"), **out)
