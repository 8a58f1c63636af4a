// oracle_body.hpp — CPU ORACLE (test infrastructure): articulated swimmer, host fp64.
//
// Restates SURVEY.md §8(a9) / Appendix A7 (1): a planar chain of spheroid links with yawing
// joints (README.md:2 "articulated underwater agent"), joint angles driven kinematically towards the
// action targets, root (centre of mass + yaw) integrated from the summed hydrodynamic wrench of the
// PREVIOUS substep with exact conservation of linear and angular momentum under shape change.
// Output per substep: marker positions / velocities rounded to fp32 once.
//
// Written independently of gym-fish_b200/csrc/body.cpp (the product's integrator); the two are
// compared through the C ABI in tests/.
#pragma once
#include "../include/fishgym.h"

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

namespace obody {

struct V2 { double x = 0, z = 0; };   // planar vectors live in the x-z plane; yaw is about +y

inline V2 axis(double th) { return {std::sin(th), std::cos(th)}; }          // tail-ward body axis
inline V2 ycross(const V2 &r) { return {r.z, -r.x}; }                       // (y_hat x r) in the plane
inline double cross_y(const V2 &a, const V2 &b) { return a.z * b.x - a.x * b.z; }   // (a x b) . y_hat

class Fish {
public:
    // n direct-forcing passes: the forcing's period-2 memory (it aims at u* + F/2 but leaves u* + F) is damped by 2 (1-l)^n only, so the
    // virtual mass that keeps the explicit body update stable grows like (1 - q) / q, q = (1-l)^n; with l = 1/2: 2^n - 1
    void forcing_passes(int n) { npass_ = n > 1 ? std::pow(2.0, std::min(n, 16)) - 1.0 : 1.0; }
    bool init(const FgFishDesc &d, std::string *why) {
        if (d.n_links < 1 || d.n_links > 8) { *why = "fish: n_links must be 1..8"; return false; }
        desc_ = d;
        nl_ = d.n_links;
        mass_.resize(nl_); inertia_.resize(nl_); first_.resize(nl_ + 1);
        cmark_.assign(nl_, 0.0); cmark2_.assign(nl_, 0.0); ctot_ = 0;
        xi_.clear(); dV_.clear();
        const double PI = 3.14159265358979323846;
        for (int k = 0; k < nl_; ++k) {
            const double a = 0.5 * d.link_len[k], r = d.link_rad[k];
            if (!(a > 0) || !(r > 0)) { *why = "fish: link_len and link_rad must be > 0"; return false; }
            const double rho_b = d.density_ratio > 0 ? d.density_ratio : 1.0;
            mass_[k] = rho_b * 4.0 / 3.0 * PI * a * r * r;
            inertia_[k] = mass_[k] * (a * a + r * r) / 5.0;
            // surface sampling: Fibonacci lattice on the unit sphere mapped onto the spheroid
            const double p = 1.6075;
            const double area = 4.0 * PI * std::pow((std::pow(r * r, p) + 2.0 * std::pow(r * a, p)) / 3.0, 1.0 / p);
            int n = d.markers_per_link > 0 ? d.markers_per_link : int(std::lround(area));
            n = std::max(n, 8);
            first_[k] = int(dV_.size());
            const double golden = PI * (3.0 - std::sqrt(5.0));
            for (int i = 0; i < n; ++i) {
                const double sz = 1.0 - (2.0 * i + 1.0) / n;
                const double rr = std::sqrt(std::max(0.0, 1.0 - sz * sz));
                const double ph = golden * i;
                const double sx = rr * std::cos(ph), sy = rr * std::sin(ph);
                xi_.push_back(r * sx); xi_.push_back(r * sy); xi_.push_back(a * sz);
                const double stretch = std::sqrt((r * a * sx) * (r * a * sx) + (r * a * sy) * (r * a * sy) + (r * r * sz) * (r * r * sz));
                dV_.push_back(4.0 * PI / n * stretch);   // area element x unit shell thickness
                // penalty stiffness of the direct forcing seen by the body: sum 2 dV (translation), sum 2 dV |xi_planar|^2 (yaw)
                cmark_[k] += 2.0 * dV_.back();
                cmark2_[k] += 2.0 * dV_.back() * ((r * sx) * (r * sx) + (a * sz) * (a * sz));
                ctot_ += 2.0 * dV_.back();
            }
        }
        first_[nl_] = int(dV_.size());
        q_.assign(n_joints(), 0.0); qd_.assign(n_joints(), 0.0);
        c_.resize(nl_); v_.resize(nl_); th_.resize(nl_); om_.resize(nl_);
        reset();
        return true;
    }

    int n_links() const { return nl_; }
    int n_joints() const { return nl_ - 1; }
    int n_markers() const { return int(dV_.size()); }
    int obs_size() const { return 8 + 2 * n_joints(); }

    void reset() {
        std::fill(q_.begin(), q_.end(), 0.0);
        std::fill(qd_.begin(), qd_.end(), 0.0);
        th0_ = desc_.heading; om0_ = 0; P_ = V2{}; L_ = 0; dP_ = V2{}; dL_ = 0;
        // place the head centre at root_pos: com = head + com_rel
        shape(th0_);
        com_.x = desc_.root_pos[0] + com_rel_.x;
        com_.z = desc_.root_pos[2] + com_rel_.z;
        place();
    }

    // one substep: wrench6[link] = hydrodynamic (force, torque about origin3[link]) ON each link
    void advance(const float *action, const double *wrench6, const double *origin3) {
        const int nj = n_joints();
        for (int j = 0; j < nj; ++j) {
            const double target = double(action[j]) * desc_.joint_limit;
            double rate = desc_.joint_gain * (target - q_[j]);
            rate = std::min(desc_.joint_rate_max, std::max(-desc_.joint_rate_max, rate));
            double qn = q_[j] + rate;
            qn = std::min(desc_.joint_limit, std::max(-desc_.joint_limit, qn));
            qd_[j] = qn - q_[j];
            q_[j] = qn;
        }
        if (desc_.free_root) {
            V2 F{};
            double T = 0;
            for (int k = 0; k < nl_; ++k) {
                const double *w = wrench6 + 6 * k;
                F.x += w[0]; F.z += w[2];
                const V2 arm{origin3[3 * k] - com_.x, origin3[3 * k + 2] - com_.z};
                T += w[4] + cross_y(arm, V2{w[0], w[2]});
            }
            double M = 0;
            for (double m : mass_) M += m;
            // geometry at the old heading and new joints
            shape(th0_);
            double I = 0, Ls = 0, Cr = 0;
            for (int k = 0; k < nl_; ++k) {
                const double r2 = rel_[k].x * rel_[k].x + rel_[k].z * rel_[k].z;
                I += inertia_[k] + mass_[k] * r2;
                Ls += inertia_[k] * omr_[k] + mass_[k] * cross_y(rel_[k], srel_[k]);
                Cr += cmark_[k] * r2 + cmark2_[k];
            }
            // explicit coupling with the direct-forcing penalty F = 2(U_d - U*) is unstable for light bodies
            // (added-mass instability); a virtual mass Mv = beta * sum(2 dV) filters the momentum increment:
            // (M + Mv) a_new = F + Mv a_old, fixed point a = F/M, unconditionally stable for beta >= 1/4
            // n direct-forcing passes per substep (multi-direct forcing): virtual mass x (2^n - 1), see forcing_passes()
            const double Mv = kBeta * npass_ * ctot_, Iv = kBeta * npass_ * Cr;
            dP_.x = (M * F.x + Mv * dP_.x) / (M + Mv);
            dP_.z = (M * F.z + Mv * dP_.z) / (M + Mv);
            dL_ = (I * T + Iv * dL_) / (I + Iv);
            P_.x += dP_.x; P_.z += dP_.z; L_ += dL_;
            com_.x += P_.x / M; com_.z += P_.z / M;
            om0_ = (L_ - Ls) / I;
            th0_ += om0_;
        }
        shape(th0_);
        place();
    }

    // markers for the coupled step, rounded to fp32 here and nowhere else
    void emit_markers(float *X, float *U, float *dV, int32_t *link, int link_offset, double *origin3) const {
        for (int k = 0; k < nl_; ++k) {
            const V2 t = axis(th_[k]);
            const V2 e{t.z, -t.x};   // lateral axis (cos th, -sin th)
            origin3[3 * k] = c_[k].x; origin3[3 * k + 1] = desc_.root_pos[1]; origin3[3 * k + 2] = c_[k].z;
            for (int i = first_[k]; i < first_[k + 1]; ++i) {
                const double a = xi_[3 * i], b = xi_[3 * i + 1], c = xi_[3 * i + 2];
                const V2 r{a * e.x + c * t.x, a * e.z + c * t.z};
                const V2 w = ycross(r);
                X[3 * i] = float(c_[k].x + r.x);
                X[3 * i + 1] = float(desc_.root_pos[1] + b);
                X[3 * i + 2] = float(c_[k].z + r.z);
                U[3 * i] = float(v_[k].x + om_[k] * w.x);
                U[3 * i + 1] = 0.f;
                U[3 * i + 2] = float(v_[k].z + om_[k] * w.z);
                dV[i] = float(dV_[i]);
                link[i] = link_offset + k;
            }
        }
    }

    void write_obs(float *o) const {
        double M = 0;
        for (double m : mass_) M += m;
        o[0] = float(c_[0].x); o[1] = float(desc_.root_pos[1]); o[2] = float(c_[0].z);
        o[3] = float(th0_);
        o[4] = float(P_.x / M); o[5] = 0.f; o[6] = float(P_.z / M);
        o[7] = float(om0_);
        for (int j = 0; j < n_joints(); ++j) { o[8 + j] = float(q_[j]); o[8 + n_joints() + j] = float(qd_[j]); }
    }

private:
    // link centres / velocities relative to the centre of mass for root heading th0 and the current joints
    void shape(double th0) {
        rel_.assign(nl_, V2{}); srel_.assign(nl_, V2{}); omr_.assign(nl_, 0.0); thl_.assign(nl_, 0.0);
        double th = th0, omr = 0;
        V2 c{}, u{};
        double M = 0;
        V2 mc{}, mu{};
        for (int k = 0; k < nl_; ++k) {
            if (k > 0) {
                const double ap = 0.5 * desc_.link_len[k - 1], an = 0.5 * desc_.link_len[k];
                const V2 tp = axis(th);
                const V2 wp = ycross(V2{ap * tp.x, ap * tp.z});
                c.x += ap * tp.x; c.z += ap * tp.z;
                u.x += omr * wp.x; u.z += omr * wp.z;
                th += q_[k - 1]; omr += qd_[k - 1];
                const V2 tn = axis(th);
                const V2 wn = ycross(V2{an * tn.x, an * tn.z});
                c.x += an * tn.x; c.z += an * tn.z;
                u.x += omr * wn.x; u.z += omr * wn.z;
            }
            thl_[k] = th; omr_[k] = omr; rel_[k] = c; srel_[k] = u;
            M += mass_[k];
            mc.x += mass_[k] * c.x; mc.z += mass_[k] * c.z;
            mu.x += mass_[k] * u.x; mu.z += mass_[k] * u.z;
        }
        com_rel_ = V2{mc.x / M, mc.z / M};
        for (int k = 0; k < nl_; ++k) {
            rel_[k].x -= com_rel_.x; rel_[k].z -= com_rel_.z;
            srel_[k].x -= mu.x / M; srel_[k].z -= mu.z / M;
        }
    }
    void place() {
        double M = 0;
        for (double m : mass_) M += m;
        V2 vc{P_.x / M, P_.z / M};
        if (!desc_.free_root) {
            // pinned root: the head centre stays at root_pos with zero velocity
            com_ = V2{desc_.root_pos[0] + com_rel_.x, desc_.root_pos[2] + com_rel_.z};
            vc = V2{-srel_[0].x, -srel_[0].z};
        }
        for (int k = 0; k < nl_; ++k) {
            c_[k] = V2{com_.x + rel_[k].x, com_.z + rel_[k].z};
            const V2 w = ycross(rel_[k]);
            v_[k] = V2{vc.x + om0_ * w.x + srel_[k].x, vc.z + om0_ * w.z + srel_[k].z};
            th_[k] = thl_[k];
            om_[k] = om0_ + omr_[k];
        }
    }

    FgFishDesc desc_{};
    int nl_ = 0;
    std::vector<double> mass_, inertia_;
    std::vector<int> first_;
    std::vector<double> xi_, dV_;          // body-frame marker points [n][3] (lateral, y, axial), volumes
    std::vector<double> q_, qd_;
    static constexpr double kBeta = 0.5;
    double npass_ = 1.0;
    std::vector<double> cmark_, cmark2_;   // per link: sum 2 dV, sum 2 dV |xi|^2 in the swimming plane
    double ctot_ = 0;
    double th0_ = 0, om0_ = 0, L_ = 0, dL_ = 0;
    V2 P_{}, dP_{}, com_{}, com_rel_{};
    std::vector<V2> rel_, srel_, c_, v_;
    std::vector<double> omr_, thl_, th_, om_;
};

}  // namespace obody
