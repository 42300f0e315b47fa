// algoplonk.hpp -- compiled-language mirror of AlgoPlonk's public API for the proving path, over the C ABI
// of include/b200plonk.h (header only; needs libb200plonk.so at link time and a B200 at run time).
//
// The reference is Go (package algoplonk) and no Go toolchain exists in the build image, so the binding a
// maintainer adds is shipped as source (go/gpuplonk).  This header is the same surface in C++, with the
// reference's names, argument meaning and error behaviour:
//
//   Compile(circuit, curve, setup)            /root/reference/algoplonk.go:37-59
//   CompiledCircuit::Verify(assignment)       /root/reference/algoplonk.go:79-98   (witness -> Prove -> verify)
//   VerifiedProof::ExportProofAndPublicInputs /root/reference/algoplonk.go:103-131
//   MarshalProof / MarshalPublicInputs        /root/reference/helper.go:13-24,91-110
//   setup::Name                               /root/reference/setup/setup.go:23-36
//   CompiledCircuit::VerifyFromInputs(inputs)   the same with the witness solved by the library (b2p_solver_*),
//                                               L R O staying in HBM (algoplonk.go:81-89: NewWitness + spr.Solve)
//   SaveKey / Compile(..., snapshot)            /root/reference/utils/utils.go:97-157 (persisted proving key)
//
// The circuit front end (gnark's frontend.Compile + NewTrace) is restated minimally: public rows first,
// last-seen-position permutation, padding rows on variable 0 -- enough for the circuits of the reference's
// examples and tests (BasicCircuit, the squaring chain).  Field values are held in Montgomery form with the
// library's own host arithmetic (csrc/field.cuh compiles for the host), i.e. in gnark's memory layout.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200plonk.h"
#include "../csrc/field.cuh"

namespace algoplonk {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc, const char* what) {
    if (rc != B2P_OK) throw Error(std::string(what) + ": " + b2p_last_error());
}

namespace setup {
// setup/setup.go:23-36
enum class Name { TestOnlyBN254 = 0, TestOnlyBLS12381, PerpetualPowersOfTauBN254, EthereumKzgCeremonyBLS12381, DuskBLS12381 };
inline bool trusted(Name n) { return n >= Name::PerpetualPowersOfTauBN254; }
inline int curve_of(Name n) { return (n == Name::TestOnlyBN254 || n == Name::PerpetualPowersOfTauBN254) ? B2P_BN254 : B2P_BLS12_381; }
}  // namespace setup

template <int CURVE> struct ScalarField;
template <> struct ScalarField<B2P_BN254> { using Fr = b2p::FrBn254; };
template <> struct ScalarField<B2P_BLS12_381> { using Fr = b2p::FrBls12381; };

// canonical big-endian hex (no 0x) -> Montgomery element; throws on overlong input
template <class Fr>
inline Fr fr_from_hex(const std::string& hex) {
    Fr raw = Fr::zero();
    if (hex.size() > (size_t)Fr::N * 8) throw Error("scalar does not fit the field");
    for (size_t i = 0; i < hex.size(); i++) {
        const char ch = hex[hex.size() - 1 - i];
        const int d = ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1;
        if (d < 0) throw Error("bad hex digit");
        raw.v[i / 8] |= (uint32_t)d << (4 * (i % 8));
    }
    return Fr::reduce_to_mont(raw);
}
template <class Fr>
inline Fr fr_from_u64(uint64_t x) {
    Fr r = Fr::zero();
    r.v[0] = (uint32_t)x;
    r.v[1] = (uint32_t)(x >> 32);
    return r.to_mont();
}

// ---- constraint system (gnark SparseR1CS, the subset the proving path needs) ------------------------
template <class Fr>
struct SparseR1CS {
    struct Constraint { Fr ql, qr, qm, qo, qk; uint32_t xa, xb, xc; };
    uint32_t nb_public = 0;
    std::vector<Fr> values;                 // solved witness, public variables first
    std::vector<Constraint> constraints;
    std::vector<uint32_t> input_vars;       // the variables an assignment gives (public, then secret)

    uint64_t domain_size() const {          // NextPowerOfTwo(nbConstraints + nbPublic), setup.go:113
        uint64_t need = constraints.size() + nb_public, n = 1;
        while (n < need) n <<= 1;
        return n < 2 ? 2 : n;
    }
};

// Eager builder: variables carry their values, one pass yields the system and its witness
template <class Fr>
struct Builder {
    SparseR1CS<Fr> cs;
    bool secret_started = false;
    uint32_t Public(const Fr& v) {
        if (secret_started) throw Error("public variables first (gnark witness order)");
        cs.values.push_back(v);
        cs.input_vars.push_back(cs.nb_public);
        return cs.nb_public++;
    }
    uint32_t Secret(const Fr& v) {
        secret_started = true;
        cs.values.push_back(v);
        cs.input_vars.push_back((uint32_t)cs.values.size() - 1);
        return (uint32_t)cs.values.size() - 1;
    }
    // a variable the solver determines from a row (not part of the assignment)
    uint32_t Internal(const Fr& v) { secret_started = true; cs.values.push_back(v); return (uint32_t)cs.values.size() - 1; }
    void Constrain(const Fr& ql, const Fr& qr, const Fr& qm, const Fr& qo, const Fr& qk, uint32_t a, uint32_t b, uint32_t c) {
        cs.constraints.push_back({ql, qr, qm, qo, qk, a, b, c});
    }
    uint32_t Mul(uint32_t a, uint32_t b) {
        const uint32_t c = Internal(cs.values[a] * cs.values[b]);
        Constrain(Fr::zero(), Fr::zero(), Fr::one(), Fr::one().neg(), Fr::zero(), a, b, c);
        return c;
    }
    uint32_t Add(uint32_t a, uint32_t b) {
        const uint32_t c = Internal(cs.values[a] + cs.values[b]);
        Constrain(Fr::one(), Fr::one(), Fr::zero(), Fr::one().neg(), Fr::zero(), a, b, c);
        return c;
    }
    void AssertIsEqual(uint32_t a, uint32_t b) {
        Constrain(Fr::one(), Fr::one().neg(), Fr::zero(), Fr::zero(), Fr::zero(), a, b, 0);
    }
};

// ---- proofs -----------------------------------------------------------------------------------------
template <int CURVE>
struct VerifiedProof {                      // algoplonk.go:28-31
    using Fr = typename ScalarField<CURVE>::Fr;
    std::vector<uint8_t> raw;               // plonk.Proof in gnark memory layout (b2p_proof_raw_size bytes)
    std::vector<Fr> witness;                // public inputs

    std::vector<uint8_t> MarshalProof() const {                      // helper.go:13-24
        std::vector<uint8_t> out(b2p_proof_marshal_size(CURVE, 0));
        check(b2p_marshal_proof(CURVE, 0, raw.data(), nullptr, out.data()), "MarshalProof");
        return out;
    }
    std::vector<uint8_t> MarshalPublicInputs() const {               // helper.go:91-110
        std::vector<uint8_t> out(32 * witness.size());
        check(b2p_marshal_public_inputs(CURVE, witness.data(), (uint32_t)witness.size(), out.data()), "MarshalPublicInputs");
        return out;
    }
    void ExportProofAndPublicInputs(const std::string& proof_path, const std::string& public_inputs_path) const {
        const auto p = MarshalProof(), w = MarshalPublicInputs();    // algoplonk.go:103-131
        std::ofstream(proof_path, std::ios::binary).write((const char*)p.data(), (std::streamsize)p.size());
        std::ofstream(public_inputs_path, std::ios::binary).write((const char*)w.data(), (std::streamsize)w.size());
    }
};

// ---- compiled circuit (algoplonk.go:21-26: Ccs + Pk + Vk + Curve; Pk resident on the GPU) ------------
template <int CURVE>
class CompiledCircuit {
public:
    using Fr = typename ScalarField<CURVE>::Fr;
    SparseR1CS<Fr> ccs;
    uint64_t n = 0;

    CompiledCircuit() = default;
    CompiledCircuit(const CompiledCircuit&) = delete;
    CompiledCircuit& operator=(const CompiledCircuit&) = delete;
    ~CompiledCircuit() {
        if (solver_) b2p_solver_free(solver_);
        if (circuit_) b2p_circuit_free(circuit_);
        if (srs_) b2p_srs_free(srs_);
    }

    // (*CompiledCircuit).Verify as the reference runs it (algoplonk.go:79-98): `inputs` are the values of the circuit's
    // public and secret variables in declaration order; the library solves the rest (b2p_solver_solve_dev: a wide
    // circuit on the GPU, a dependency chain on a host thread), proves from HBM (b2p_prove_dev) and verifies.
    // Hint-free circuits only in this mirror (its Builder has no hints); throws "constraint #i is not satisfied" like gnark.
    VerifiedProof<CURVE> VerifyFromInputs(const std::vector<Fr>& inputs, const std::vector<Fr>& blinding, bool self_check = true) {
        if (blinding.size() != 9) throw Error("9 blinding scalars expected");
        if (inputs.size() != ccs.input_vars.size()) throw Error("one value per public / secret variable expected");
        if (!solver_) {
            std::vector<uint32_t> xa(n, 0), xb(n, 0), xc(n, 0);
            for (uint32_t i = 0; i < ccs.nb_public; i++) xa[i] = i;
            for (size_t j = 0; j < ccs.constraints.size(); j++) {
                xa[ccs.nb_public + j] = ccs.constraints[j].xa;
                xb[ccs.nb_public + j] = ccs.constraints[j].xb;
                xc[ccs.nb_public + j] = ccs.constraints[j].xc;
            }
            check(b2p_solver_create(CURVE, n, ccs.nb_public, ccs.values.size(), ccs.input_vars.data(),
                                    (uint32_t)ccs.input_vars.size(), ql_.data(), qr_.data(), qm_.data(), qo_.data(), qk_.data(),
                                    xa.data(), xb.data(), xc.data(), &solver_), "solver");
        }
        void *dL = nullptr, *dR = nullptr, *dO = nullptr;
        check(b2p_solver_solve_dev(solver_, inputs.data(), B2P_SOLVE_AUTO, &dL, &dR, &dO), "spr.Solve");
        VerifiedProof<CURVE> vp;
        vp.raw.resize(b2p_proof_raw_size(CURVE, 0));
        check(b2p_prove_dev(circuit_, dL, dR, dO, nullptr, nullptr, blinding.data(), vp.raw.data()), "plonk.Prove");
        vp.witness.assign(inputs.begin(), inputs.begin() + ccs.nb_public);
        if (self_check) VerifyProof(vp.MarshalProof(), vp.MarshalPublicInputs());
        return vp;
    }
    // utils.SerializeCompiledCircuit (utils/utils.go:97-121) for the circuit half of the key: the library's snapshot,
    // which Compile(..., snapshot_path) loads back instead of rebuilding the trace
    void SaveKey(const std::string& path) const {
        check(b2p_circuit_save(path.c_str(), CURVE, n, ccs.nb_public, ql_.data(), qr_.data(), qm_.data(), qo_.data(),
                               qk_.data(), perm_.data(), 0, nullptr, nullptr, nullptr, 0), "SerializeCompiledCircuit");
    }

    // plonk.Prove alone (algoplonk.go:89 without :93): the explicit opt-out of the self-check below
    VerifiedProof<CURVE> ProveOnly(const std::vector<Fr>& blinding) const {
        return Verify(blinding, [](const std::vector<uint8_t>&, const std::vector<uint8_t>&) { return true; }, false);
    }
    // algoplonk.go:79-98: witness -> Prove -> plonk.Verify (+ `verifier`, if given).
    // This mirror covers circuits WITHOUT BSB22 commitments (k = 0 in b2p_prove / b2p_verify / MarshalProof below):
    // its Builder has no Commit; circuits with commitments go through the C ABI directly (or the Python / Go hosts).
    // `blinding`: the 9 scalars gnark draws with fr.SetRandom (an input, so proofs are reproducible).
    template <class Verifier = std::nullptr_t>
    VerifiedProof<CURVE> Verify(const std::vector<Fr>& blinding, Verifier verifier = nullptr, bool self_check = true) const {
        if (blinding.size() != 9) throw Error("9 blinding scalars expected");
        std::vector<Fr> L(n, ccs.values.empty() ? Fr::zero() : ccs.values[0]), R = L, O = L;   // padding rows: variable 0
        for (uint32_t i = 0; i < ccs.nb_public; i++) L[i] = ccs.values[i];
        for (size_t j = 0; j < ccs.constraints.size(); j++) {
            const auto& c = ccs.constraints[j];
            L[ccs.nb_public + j] = ccs.values[c.xa];
            R[ccs.nb_public + j] = ccs.values[c.xb];
            O[ccs.nb_public + j] = ccs.values[c.xc];
        }
        VerifiedProof<CURVE> vp;
        vp.raw.resize(b2p_proof_raw_size(CURVE, 0));
        check(b2p_prove(circuit_, L.data(), R.data(), O.data(), nullptr, nullptr, blinding.data(), vp.raw.data()), "plonk.Prove");
        vp.witness.assign(ccs.values.begin(), ccs.values.begin() + ccs.nb_public);
        // plonk.Verify, algoplonk.go:93: the reference always verifies.  Without the setup's G2 points and without
        // a verifier callback nothing could check the proof: that is an error, not a silent pass (ProveOnly() is
        // the explicit opt-out).
        if (self_check && kzg_g2_.empty() && std::is_same<Verifier, std::nullptr_t>::value)
            throw Error("error verifying proof: the setup's G2 points are unknown (LoadKzgVk / SetKzgG2), "
                        "or call ProveOnly()");
        if (self_check && !kzg_g2_.empty()) VerifyProof(vp.MarshalProof(), vp.MarshalPublicInputs());
        if constexpr (!std::is_same<Verifier, std::nullptr_t>::value) {
            if (!verifier(vp.MarshalProof(), vp.MarshalPublicInputs())) throw Error("error verifying proof");
        }
        return vp;
    }
    // plonk.Verify(proof, cc.Vk, publicWitness) on marshalled bytes (b2p_verify: host arithmetic of the library);
    // throws "error verifying proof: <failed check>" when the proof is rejected
    void VerifyProof(const std::vector<uint8_t>& proof, const std::vector<uint8_t>& public_inputs) const {
        if (kzg_g2_.empty()) throw Error("the setup's G2 points are unknown (SetKzgG2)");
        const auto vk = VkCommitments();
        std::vector<uint8_t> g1(CURVE == B2P_BN254 ? 64 : 96);
        check(b2p_srs_get_points(srs_, 0, 1, g1.data()), "vk.Kzg.G1");
        check(b2p_verify(CURVE, n, ccs.nb_public, 0, nullptr, vk.data(), g1.data(), kzg_g2_.data(), proof.data(),
                         proof.size(), public_inputs.empty() ? nullptr : public_inputs.data(), public_inputs.size()),
              "plonk.Verify");
    }
    // vk.Kzg.G2[0], vk.Kzg.G2[1] of a trusted setup (two G2Affine in gnark memory layout, from its vk.bin);
    // the TestOnly setups derive theirs from tau in Compile
    void SetKzgG2(const void* two_g2_affine) {
        const auto* p = static_cast<const uint8_t*>(two_g2_affine);
        kzg_g2_.assign(p, p + 8 * (CURVE == B2P_BN254 ? 32 : 48));
    }
    // the same from the bytes of the embedded setup/<name>/vk.bin (srs.Vk.ReadFrom, setup/setup.go:174,190)
    void LoadKzgVk(const void* vk_bin, uint64_t vk_len) {
        std::vector<uint8_t> g2(8 * (CURVE == B2P_BN254 ? 32 : 48)), g1(2 * (CURVE == B2P_BN254 ? 32 : 48));
        check(b2p_kzg_vk_load(CURVE, vk_bin, vk_len, g2.data(), g1.data()), "srs.Vk.ReadFrom");
        kzg_g2_ = g2;
    }
    // S1 S2 S3 Ql Qr Qm Qo Qk commitments of the verifying key (G1Affine memory layout)
    std::vector<uint8_t> VkCommitments() const {
        std::vector<uint8_t> out(8 * (CURVE == B2P_BN254 ? 64 : 96));
        check(b2p_circuit_vk_commitments(circuit_, out.data()), "vk commitments");
        return out;
    }

    template <int C> friend CompiledCircuit<C>* CompileInto(CompiledCircuit<C>*, const SparseR1CS<typename ScalarField<C>::Fr>&,
                                                           setup::Name, const typename ScalarField<C>::Fr*, const void*, uint64_t,
                                                           const char*);
private:
    b2p_srs* srs_ = nullptr;
    b2p_circuit* circuit_ = nullptr;
    b2p_solver* solver_ = nullptr;
    std::vector<uint8_t> kzg_g2_;
    std::vector<Fr> ql_, qr_, qm_, qo_, qk_;     // the trace (the solver's rows, SaveKey)
    std::vector<int64_t> perm_;
};

template <int CURVE>
inline CompiledCircuit<CURVE>* CompileInto(CompiledCircuit<CURVE>* cc, const SparseR1CS<typename ScalarField<CURVE>::Fr>& cs,
                                           setup::Name name, const typename ScalarField<CURVE>::Fr* test_tau,
                                           const void* pk_bin, uint64_t pk_len, const char* snapshot_path = nullptr) {
    using Fr = typename ScalarField<CURVE>::Fr;
    if ((int)name < 0 || (int)name > (int)setup::Name::DuskBLS12381) throw Error("unknown setup");   // compile_test.go:22-30
    if (setup::curve_of(name) != CURVE) throw Error("curve and trusted setup do not match");          // algoplonk.go:46-48
    check(b2p_init(-1), "b2p_init");
    cc->ccs = cs;
    const uint64_t n = cc->n = cs.domain_size();
    if (setup::trusted(name)) {
        if (!pk_bin) throw Error("trusted setups need their pk.bin bytes");
        check(b2p_srs_load_compressed(CURVE, pk_bin, pk_len, n + 3, &cc->srs_), "setup.Run");          // setup.go:113-139
    } else {
        if (!test_tau) throw Error("TestOnly setups need a tau");
        check(b2p_srs_generate_unsafe(CURVE, test_tau, n + 3, &cc->srs_), "unsafekzg.NewSRS");         // setup.go:102-108
        cc->kzg_g2_.resize(8 * (CURVE == B2P_BN254 ? 32 : 48));
        check(b2p_g2_generate_unsafe(CURVE, test_tau, cc->kzg_g2_.data()), "unsafekzg.NewSRS (G2)");
    }
    // trace: Lagrange columns + permutation (gnark NewTrace / buildPermutation)
    std::vector<Fr> ql(n, Fr::zero()), qr = ql, qm = ql, qo = ql, qk = ql;
    const uint32_t off = cs.nb_public;
    for (uint32_t i = 0; i < off; i++) ql[i] = Fr::one().neg();
    std::vector<int64_t> lro(3 * n, 0), perm(3 * n, -1);
    for (uint32_t i = 0; i < off; i++) lro[i] = i;
    for (size_t j = 0; j < cs.constraints.size(); j++) {
        const auto& c = cs.constraints[j];
        ql[off + j] = c.ql; qr[off + j] = c.qr; qm[off + j] = c.qm; qo[off + j] = c.qo; qk[off + j] = c.qk;
        lro[off + j] = c.xa; lro[n + off + j] = c.xb; lro[2 * n + off + j] = c.xc;
    }
    std::vector<int64_t> cycle(cs.values.empty() ? 1 : cs.values.size(), -1);
    for (uint64_t i = 0; i < 3 * n; i++) {
        const int64_t v = lro[i];
        if (cycle[v] != -1) perm[i] = cycle[v];
        cycle[v] = (int64_t)i;
    }
    for (uint64_t i = 0; i < 3 * n; i++)
        if (perm[i] == -1) perm[i] = cycle[lro[i]];
    // utils.DeserializeCompiledCircuit (utils/utils.go:124-157): a snapshot written by SaveKey goes from the page cache
    // to HBM; without one (or if it does not load) the columns built above are uploaded
    if (!snapshot_path || b2p_circuit_load_file(cc->srs_, snapshot_path, &cc->circuit_) != B2P_OK)
        check(b2p_circuit_load(cc->srs_, n, cs.nb_public, ql.data(), qr.data(), qm.data(), qo.data(), qk.data(), perm.data(), 0,
                               nullptr, nullptr, nullptr, 0, &cc->circuit_), "plonk.Setup");
    cc->ql_ = std::move(ql); cc->qr_ = std::move(qr); cc->qm_ = std::move(qm); cc->qo_ = std::move(qo); cc->qk_ = std::move(qk);
    cc->perm_ = std::move(perm);
    return cc;
}

// algoplonk.go:37-59.  TestOnly setups: `test_tau` (unsafekzg draws a random one); trusted setups: the bytes of
// the embedded setup/<name>/pk.bin.
template <int CURVE>
inline void Compile(CompiledCircuit<CURVE>& out, const SparseR1CS<typename ScalarField<CURVE>::Fr>& cs, setup::Name name,
                    const typename ScalarField<CURVE>::Fr* test_tau = nullptr, const void* pk_bin = nullptr, uint64_t pk_len = 0,
                    const char* snapshot_path = nullptr) {
    CompileInto<CURVE>(&out, cs, name, test_tau, pk_bin, pk_len, snapshot_path);
}

}  // namespace algoplonk
