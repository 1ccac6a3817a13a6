// sx_emul.cpp -- TEST-ONLY host harness: runs the product's per-window automaton
// (stringsext_b200/csrc/sx_core.cuh) sequentially on the CPU so that the window decomposition,
// transfer-function classification and emit rules can be differential-tested against the oracle
// without a GPU.  It is NOT part of the product library and nothing in stringsext_b200/ loads it.
#include "../../stringsext_b200/csrc/sx_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace sx;

struct HostTile {
    GlobalSrc g;
    uint8_t get(int64_t off) const { return g.get(off); }
    template <class F> void for_each_byte(int64_t ws, int64_t we, F&& f) const {
        for (int64_t p = ws; p < we; ++p) f((uint32_t)g.get(p), p);
    }
};

template <class Dec>
static void run(const ScanParams& P, std::vector<Record>& recs, std::vector<uint8_t>& text, Carry* final_carry,
                int32_t* final_npend, uint64_t* stats) {
    GlobalSrc g{P.in, P.pend};
    HostTile ts{g};
    Geometry geo;
    geo.init(P);
    const int64_t nslices = (P.len + P.slice_len - 1) / P.slice_len;
    const int64_t nwin_max = nslices * geo.wps;
    std::vector<WinDesc> desc;
    std::vector<WinGeom> geos;
    std::vector<int32_t> npend;
    for (int64_t w = 0; w < nwin_max; ++w) {
        WinGeom wg;
        if (!geo.window(w, wg)) continue;
        WinResult r;
        WinDesc d;
        scan_window<Dec>(P, ts, g, wg, carry_none(), MODE_COUNT, nullptr, 0, r, &d);
        desc.push_back(d);
        geos.push_back(wg);
        npend.push_back(r.npend_out);
    }
    const size_t nw = desc.size();
    std::vector<Carry> kin(nw + 1);
    kin[0] = P.k0;
    for (size_t i = 0; i < nw; ++i) {
        const WinDesc& d = desc[i];
        if (d.type == WT_CONST) kin[i + 1] = d.null_out;
        else if (d.type == WT_CASEB) kin[i + 1] = eval_caseb(P, d, kin[i], (uint32_t)(geos[i].we - geos[i].ws));
        else {
            WinResult r;
            scan_window<Dec>(P, ts, g, geos[i], kin[i], MODE_STATE, nullptr, 0, r, nullptr);
            kin[i + 1] = r.out;
            stats[1]++;
        }
        stats[0]++;
        if (d.type == WT_CASEB) stats[2]++;
        // self-check: the classification must agree with a replay
        {
            WinResult r;
            scan_window<Dec>(P, ts, g, geos[i], kin[i], MODE_STATE, nullptr, 0, r, nullptr);
            if (memcmp(&r.out, &kin[i + 1], sizeof(Carry)) != 0) {
                stats[3]++;
                kin[i + 1] = r.out;
            }
        }
    }
    for (size_t i = 0; i < nw; ++i) {
        const bool e = needs_emit(P, desc[i], kin[i]);
        WinResult rc;
        scan_window<Dec>(P, ts, g, geos[i], kin[i], MODE_COUNT, nullptr, 0, rc, nullptr);
        if (!e) {
            if (rc.nrec != 0) stats[4]++;  // emit rule missed a yielding window
            continue;
        }
        if (carry_is_null(kin[i]) && desc[i].nrec != 0xFFFF && (rc.nrec != desc[i].nrec || rc.ntext != desc[i].ntext)) stats[5]++;
        const size_t base = recs.size();
        recs.resize(base + rc.nrec);
        WinResult rw;
        scan_window<Dec>(P, ts, g, geos[i], kin[i], MODE_WRITE, recs.data() + base, text.size(), rw, nullptr);
        const size_t tb = text.size();
        text.resize(tb + rc.ntext + 8);
        for (size_t k = base; k < recs.size(); ++k) {
            Record& r = recs[k];
            const uint32_t n = transcode_range(P, g, r.in_start, r.in_len, text.data() + r.text_off);
            if (n != r.text_len) stats[6]++;
        }
        text.resize(tb + rc.ntext);
    }
    *final_carry = nw ? kin[nw] : P.k0;
    *final_npend = nw ? npend[nw - 1] : P.npend;
    if (final_carry->kind == K_L && final_carry->k > 0) {
        Record r{};
        r.in_start = P.len - (int64_t)final_carry->in_bytes;
        r.in_len = final_carry->in_bytes - (uint32_t)*final_npend;
        r.text_len = final_carry->out_bytes;
        r.text_off = text.size();
        r.flags = RF_LEFTOVER | ((final_carry->flags & CF_HOSTCARRY) ? RF_HOSTCARRY : 0);
        text.resize(text.size() + r.text_len + 8);
        transcode_range(P, g, r.in_start, r.in_len, text.data() + r.text_off);
        text.resize(text.size() - 8);
        recs.push_back(r);
    }
}

extern "C" {
struct emul_out {
    Record* recs;
    size_t nrecs;
    uint8_t* text;
    size_t ntext;
    Carry final_carry;
    int32_t final_npend;
    uint64_t stats[8];
};

// Returns 0 on success.  `params` is a fully populated ScanParams (in = host pointer).
int sx_emul_scan(const ScanParams* P, emul_out* out) {
    std::vector<Record> recs;
    std::vector<uint8_t> text;
    memset(out->stats, 0, sizeof out->stats);
    switch (P->enc) {
    case ENC_XUD: run<DecXud>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_UTF8: run<DecUtf8>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_UTF16LE: run<DecUtf16<false>>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_UTF16BE: run<DecUtf16<true>>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_SB: run<DecSb>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_UTF32LE: run<DecUtf32<false>>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    case ENC_UTF32BE: run<DecUtf32<true>>(*P, recs, text, &out->final_carry, &out->final_npend, out->stats); break;
    default: return 1;
    }
    out->nrecs = recs.size();
    out->recs = (Record*)malloc(sizeof(Record) * (recs.size() + 1));
    memcpy(out->recs, recs.data(), sizeof(Record) * recs.size());
    out->ntext = text.size();
    out->text = (uint8_t*)malloc(text.size() + 1);
    memcpy(out->text, text.data(), text.size());
    return 0;
}
void sx_emul_free(emul_out* o) { free(o->recs); free(o->text); }
size_t sx_emul_sizeof_params() { return sizeof(ScanParams); }
}
