// Drives the C++ shim (ldso_b200/host/ldso_shim.hpp) the way FullSystem::optimize drives the reference classes
// (src/frontend/FullSystem.cc:725-864 restricted to the path) and prints the observables as JSON for the pytest wrapper.
//   shim_test <window.bin> [fused]
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../../ldso_b200/host/ldso_shim.hpp"

using namespace ldso;
using namespace ldso::internal;

static void rd(FILE *f, void *p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const bool fused = argc > 2 && std::string(argv[2]) == "fused";
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[6];   // w h levels nF nP nR
    rd(f, hdr, sizeof(hdr));
    const int w = hdr[0], h = hdr[1], levels = hdr[2], nF = hdr[3], nP = hdr[4], nR = hdr[5];
    double K[4];
    rd(f, K, sizeof(K));
    auto HCalib = std::make_shared<CalibHessian>(K[0], K[1], K[2], K[3]);
    auto ef = std::make_shared<EnergyFunctional>(w, h, levels);
    if (!ef->ok()) { printf("{\"error\": \"no device\"}\n"); return 1; }
    std::vector<std::vector<std::vector<float>>> pyr(nF);
    std::vector<shared_ptr<FrameHessian>> frameHessians;
    for (int i = 0; i < nF; i++) {
        auto fr = std::make_shared<Frame>();
        auto fh = std::make_shared<FrameHessian>(fr);
        SE3 ev;
        rd(f, ev.R, 72); rd(f, ev.t, 24);
        Vec10 sz, st;
        rd(f, sz.v, 80); rd(f, st.v, 80);
        float ab; int fid;
        rd(f, &ab, 4); rd(f, &fid, 4);
        fr->id = fid; fh->frameID = i; fh->ab_exposure = ab;
        fh->setEvalPT(ev, sz);
        fh->setState(st);
        pyr[i].resize(levels);
        for (int l = 0; l < levels; l++) {
            pyr[i][l].resize((size_t) 3 * (w >> l) * (h >> l));
            rd(f, pyr[i][l].data(), 4 * pyr[i][l].size());
            fh->dIp[l] = pyr[i][l].data();
        }
        frameHessians.push_back(fh);
        ef->insertFrame(fh, HCalib);
    }
    std::vector<int> host(nP), rb(nP + 1), tgt(nR);
    std::vector<float> u(nP), v(nP), id(nP), idz(nP), col(8 * (size_t) nP), wts(8 * (size_t) nP);
    rd(f, host.data(), 4 * nP); rd(f, u.data(), 4 * nP); rd(f, v.data(), 4 * nP); rd(f, id.data(), 4 * nP); rd(f, idz.data(), 4 * nP);
    rd(f, col.data(), 32 * (size_t) nP); rd(f, wts.data(), 32 * (size_t) nP); rd(f, rb.data(), 4 * (nP + 1)); rd(f, tgt.data(), 4 * nR);
    fclose(f);
    std::vector<shared_ptr<PointFrameResidual>> activeResiduals;
    for (int p = 0; p < nP; p++) {
        auto ph = std::make_shared<PointHessian>();
        ph->u = u[p]; ph->v = v[p];
        ph->setIdepthZero(idz[p]); ph->setIdepth(id[p]);
        memcpy(ph->color, &col[8 * (size_t) p], 32); memcpy(ph->weights, &wts[8 * (size_t) p], 32);
        ph->hostFrame = frameHessians[host[p]];
        frameHessians[host[p]]->pointHessians.push_back(ph);
        ef->nPoints++;
        for (int r = rb[p]; r < rb[p + 1]; r++) {
            auto res = std::make_shared<PointFrameResidual>(ph, frameHessians[host[p]], frameHessians[tgt[r]]);
            ph->residuals.push_back(res);
            ef->insertResidual(res);
            activeResiduals.push_back(res);
        }
    }
    ef->makeIDX();

    auto linearizeAll = [&]() {     // FullSystem::linearizeAll(false)
        double E = 0;
        for (auto &r : activeResiduals) E += r->linearize(HCalib);
        return E;
    };
    printf("{");
    if (!fused) {
        for (auto &r : activeResiduals) r->resetOOB();
        double E = linearizeAll();
        for (auto &r : activeResiduals) r->applyRes(true);
        printf("\"energy\": [%.9g", E);
        std::vector<double> xs;
        for (int it = 0; it < 3; it++) {
            // backupState (FullSystem.cc:1662-1676)
            HCalib->value_backup = HCalib->value;
            for (auto &fh : frameHessians) fh->state_backup = fh->get_state();
            for (auto &fh : frameHessians) for (auto &ph : fh->pointHessians) ph->idepth_backup = ph->idepth;
            ef->solveSystemF(it, 1e-1, HCalib);
            if (it == 0) xs.assign(ef->lastX.d.begin(), ef->lastX.d.end());
            // doStepFromBackup(1,1,1,1,1) (FullSystem.cc:1587-1622)
            VecC nv; for (int i = 0; i < 4; i++) nv[i] = HCalib->value_backup[i] + HCalib->step[i];
            HCalib->setValue(nv);
            for (auto &fh : frameHessians) { Vec10 ns; for (int i = 0; i < 10; i++) ns[i] = fh->state_backup[i] + fh->step[i]; fh->setState(ns); }
            for (auto &fh : frameHessians) for (auto &ph : fh->pointHessians) { ph->setIdepth(ph->idepth_backup + ph->step); ph->setIdepthZero(ph->idepth_backup + ph->step); }
            ef->setDeltaF(HCalib);
            E = linearizeAll();
            for (auto &r : activeResiduals) r->applyRes(true);
            printf(", %.9g", E);
        }
        printf("], \"resInA\": %d, \"lastX0\": [", ef->resInA);
        for (size_t i = 0; i < xs.size(); i++) printf("%s%.12g", i ? ", " : "", xs[i]);
        printf("]");
    } else {
        bool ok = ef->optimizeOnDevice(0, 3, HCalib);
        printf("\"ok\": %s, \"energy\": [%.9g]", ok ? "true" : "false", ef->lastEnergy);
    }
    if (argc > 2 && std::string(argv[2]) == "margframe") {
        // EnergyFunctional::marginalizeFrame on a synthetic prior: HM = diag(50 + 3i) + v v^T, v_i = 20 sin(i+1), bM_i = 10 cos(i)
        const int n = 8 * nF + 4;
        ef->HM.resize(n, n); ef->bM.d.assign(n, 0.0);
        for (int j = 0; j < n; j++) {
            ef->bM.d[j] = 10.0 * std::cos((double) j);
            for (int i = 0; i < n; i++) ef->HM(i, j) = 20.0 * std::sin(i + 1.0) * 20.0 * std::sin(j + 1.0) + (i == j ? 50.0 + 3.0 * i : 0.0);
        }
        ldso::internal::stateEpoch()++;
        // the frame's points/residuals leave the window first (FullSystem::marginalizeFrame drops them)
        auto victim = frameHessians[1];
        for (auto &fh : frameHessians) for (auto &ph : fh->pointHessians) {
            std::vector<shared_ptr<PointFrameResidual>> keep;
            for (auto &r : ph->residuals) if (r->target.lock() != victim && r->host.lock() != victim) keep.push_back(r);
            ph->residuals = keep;
        }
        victim->pointHessians.clear();
        bool ok = ef->marginalizeFrame(victim);
        printf(", \"marg_ok\": %s, \"marg_n\": %d, \"marg_nframes\": %d, \"HM\": [", ok ? "true" : "false", ef->HM.r, ef->nFrames);
        for (size_t i = 0; i < ef->HM.d.size(); i++) printf("%s%.15g", i ? ", " : "", ef->HM.d[i]);
        printf("], \"bM\": [");
        for (size_t i = 0; i < ef->bM.d.size(); i++) printf("%s%.15g", i ? ", " : "", ef->bM.d[i]);
        printf("]}\n");
        return 0;
    }
    int nIn = 0, nOob = 0, nAct = 0;
    for (auto &r : activeResiduals) { nIn += r->state_state == ResState::IN; nOob += r->state_state == ResState::OOB; nAct += r->isActive(); }
    printf(", \"nIn\": %d, \"nOOB\": %d, \"nActive\": %d, \"idepth\": [", nIn, nOob, nAct);
    bool first = true;
    for (auto &fh : frameHessians) for (auto &ph : fh->pointHessians) { printf("%s%.8g", first ? "" : ", ", ph->idepth); first = false; }
    printf("], \"state\": [");
    first = true;
    for (auto &fh : frameHessians) for (int i = 0; i < 8; i++) { printf("%s%.12g", first ? "" : ", ", fh->get_state()[i]); first = false; }
    printf("]");
    // the coarse tracker against the newest keyframe as reference and the oldest as the "new" frame
    {
        CoarseTracker tr(w, h, levels);
        tr.makeK(HCalib);
        tr.setCoarseTrackingRef(frameHessians);
        SE3 T;   // identity initial guess
        AffLight aff(0, 0);
        Vec5 mr; for (int i = 0; i < 5; i++) mr[i] = NAN;
        bool ok = tr.trackNewestCoarse(frameHessians[nF - 2], T, aff, levels - 1, mr);
        auto num = [](double v) { static char b[4][64]; static int k = 0; k = (k + 1) & 3; if (std::isfinite(v)) snprintf(b[k], 64, "%.9g", v); else snprintf(b[k], 64, "null"); return b[k]; };
        printf(", \"track_ok\": %s, \"track_t\": [%s, %s, %s], \"track_res\": [%s, %s]", ok ? "true" : "false", num(T.t[0]), num(T.t[1]), num(T.t[2]),
               num(tr.lastResiduals[0]), num(tr.lastResiduals[1]));
    }
    printf("}\n");
    return 0;
}
