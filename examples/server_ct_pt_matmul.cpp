// examples/server_ct_pt_matmul.cpp — the deployment split in ~60 lines.
//
//   client (stock SEAL 4.1, unchanged):  keygen; encrypt the column-packed input (batch_input);
//                                        ct.save(stream) / relin_keys.save(stream) / galois_keys.save(stream)
//                                        — normally the seeded Serializable<> forms, half the bytes
//   server (this file, B200):            load the streams through the seal:: facade, run MOAI's module code
//                                        (the reference's own headers, unmodified), save the encrypted result
//
// Build (no SEAL needed on the server):
//   g++ -std=c++17 -O2 -fopenmp -I$REPO/include/facade_fused -I$REPO/include/facade -I$REPO/include -I$MOAI/include \
//       examples/server_ct_pt_matmul.cpp -L$REPO/moai-fhe-transformerinference-public_b200 -lmoai_b200 -o server
// Usage:  server <params.txt> <inputs.seal> <weights.txt> <outputs.seal>
//   params.txt : log2(N) followed by the coeff_modulus bit sizes the client used (e.g. "16 51 46 46 ... 58")
//   inputs.seal: K ciphertext streams back to back (Ciphertext::save, compr_mode_type::none)
//   weights.txt: K, C, then K*C doubles (row-major)
#include "seal/seal.h" // include/facade/seal/seal.h -> include/moai_b200_seal.hpp

#include <fstream>
#include <iostream>
#include <vector>

#include "source/matrix_mul/Ct_pt_matrix_mul.hpp" // the reference's header name; with facade_fused: one device pipeline

using namespace seal;
using namespace std;

int main(int argc, char **argv)
{
    if (argc != 5)
    {
        cerr << "usage: " << argv[0] << " params.txt inputs.seal weights.txt outputs.seal" << endl;
        return 2;
    }
    try
    {
        // the client's parameters (primes follow from the bit sizes exactly as on the client: CoeffModulus::Create)
        ifstream pf(argv[1]);
        int log_n = 0, b = 0;
        pf >> log_n;
        vector<int> bits;
        while (pf >> b)
        {
            bits.push_back(b);
        }
        EncryptionParameters parms(scheme_type::ckks);
        parms.set_poly_modulus_degree(size_t(1) << log_n);
        parms.set_coeff_modulus(CoeffModulus::Create(size_t(1) << log_n, bits));
        SEALContext context(parms, true, sec_level_type::none); // throws without a CUDA device: no CPU fallback

        ifstream wf(argv[3]);
        int K = 0, C = 0;
        wf >> K >> C;
        vector<vector<double>> W(K, vector<double>(C));
        for (auto &row : W)
        {
            for (auto &w : row)
            {
                wf >> w;
            }
        }

        // encrypted inputs in SEAL's wire format (parms_id checked against this context's levels)
        ifstream in(argv[2], ios::binary);
        vector<Ciphertext> enc_X(K);
        for (auto &ct : enc_X)
        {
            ct.load(context, in);
        }

        // MOAI's module function, same name and signature as M/source/matrix_mul/Ct_pt_matrix_mul.hpp:4-6
        vector<Ciphertext> enc_Y = ct_pt_matrix_mul_wo_pre(enc_X, W, K, C, K, context);

        ofstream out(argv[4], ios::binary);
        for (auto &ct : enc_Y)
        {
            ct.save(out); // the client reads these with stock Ciphertext::load and decrypts
        }
        cout << C << " ciphertexts at chain_index " << context.get_context_data(enc_Y[0].parms_id())->chain_index()
             << ", scale 2^" << log2(enc_Y[0].scale()) << endl;
    }
    catch (const exception &e)
    {
        cerr << "error: " << e.what() << endl;
        return 1;
    }
    return 0;
}
