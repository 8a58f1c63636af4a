// body.hpp — articulated swimmer integrated on the host every substep (BASELINE.json:5 (c)).
//
// "physical articulated underwater agent" (/root/reference/README.md:2); no reference source exists.
// Model: planar chain of prolate-spheroid links in the x-z plane (z = swim axis), revolute joints about +y.
// Joint angles follow the action targets through a rate-limited first-order servo; the root (centre of
// mass and yaw) is integrated from the summed hydrodynamic wrench of the previous substep so that linear
// and angular momentum are conserved exactly under shape change.  Only marker positions / velocities
// (fp32, rounded once here) and the per-link wrench reductions cross PCIe.
#pragma once
#include "../../include/fishgym.h"

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

namespace fg {

class Fish {
public:
    // Multiplier of the virtual mass for n direct-forcing passes per substep.  With the Guo half-force a pass aims at u* + F/2 = U_d
    // while the fluid leaves the step with u* + F: the forcing overshoots, and its period-2 memory, F_{n+1} ~ (2 (1-l)^n - 1) F_n for
    // a mode of the interpolate-spread operator with eigenvalue l, is damped less the more passes enforce no-slip.  Coupled to the
    // explicit body update this is stable only for K < (M + 2 Mv) 2 (1-l)^n with the n-pass stiffness K = 2 sum(dV) (1 - (1-l)^n) / l,
    // i.e. the virtual mass has to grow like ((1 - q) / q) with q = (1-l)^n; l = 1/2 (an upper value for the translation and yaw modes
    // of closed surfaces at one marker per unit area: measured 0.42 on the test fish) gives 2^n - 1.
    void set_forcing_passes(int n) { passes_ = n > 1 ? std::ldexp(1.0, std::min(n, 16)) - 1.0 : 1.0; }
    bool init(const FgFishDesc &d, std::string &why) {
        if (d.n_links < 1 || d.n_links > 8) { why = "fish: n_links must be 1..8"; return false; }
        d_ = d;
        n_ = d.n_links;
        const double pi = 3.14159265358979323846;
        const double rho_b = d.density_ratio > 0 ? d.density_ratio : 1.0;
        half_.resize(n_); m_.resize(n_); I_.resize(n_); start_.assign(n_ + 1, 0);
        pts_.clear(); vol_.clear();
        pen_.assign(n_, 0.0); pen2_.assign(n_, 0.0); pen_total_ = 0;
        M_ = 0;
        for (int k = 0; k < n_; ++k) {
            const double a = 0.5 * d.link_len[k], r = d.link_rad[k];
            if (!(a > 0) || !(r > 0)) { why = "fish: link_len and link_rad must be > 0"; return false; }
            half_[k] = a;
            m_[k] = rho_b * 4.0 / 3.0 * pi * a * r * r;
            I_[k] = m_[k] * (a * a + r * r) / 5.0;          // spheroid about an axis normal to its symmetry axis
            M_ += m_[k];
            // Fibonacci lattice on the unit sphere, stretched onto the spheroid (r, r, a); one marker per unit area
            const double pw = 1.6075;
            const double area = 4.0 * pi * std::pow((std::pow(r * r, pw) + 2.0 * std::pow(r * a, pw)) / 3.0, 1.0 / pw);
            int cnt = d.markers_per_link > 0 ? d.markers_per_link : int(std::lround(area));
            cnt = std::max(cnt, 8);
            start_[k] = int(vol_.size());
            const double ga = pi * (3.0 - std::sqrt(5.0));
            for (int i = 0; i < cnt; ++i) {
                const double s3 = 1.0 - (2.0 * i + 1.0) / cnt;
                const double rr = std::sqrt(std::max(0.0, 1.0 - s3 * s3));
                const double s1 = rr * std::cos(ga * i), s2 = rr * std::sin(ga * i);
                pts_.push_back(r * s1); pts_.push_back(r * s2); pts_.push_back(a * s3);
                const double e1 = r * a * s1, e2 = r * a * s2, e3 = r * r * s3;
                vol_.push_back(4.0 * pi / cnt * std::sqrt(e1 * e1 + e2 * e2 + e3 * e3));
                pen_[k] += 2.0 * vol_.back();
                pen2_[k] += 2.0 * vol_.back() * ((r * s1) * (r * s1) + (a * s3) * (a * s3));
                pen_total_ += 2.0 * vol_.back();
            }
        }
        start_[n_] = int(vol_.size());
        const int nj = n_ - 1;
        q_.assign(nj, 0.0); qd_.assign(nj, 0.0);
        rx_.assign(n_, 0.0); rz_.assign(n_, 0.0); sx_.assign(n_, 0.0); sz_.assign(n_, 0.0);
        wrel_.assign(n_, 0.0); ang_.assign(n_, 0.0);
        cx_.assign(n_, 0.0); cz_.assign(n_, 0.0); vx_.assign(n_, 0.0); vz_.assign(n_, 0.0); w_.assign(n_, 0.0);
        reset();
        return true;
    }

    int n_links() const { return n_; }
    int n_joints() const { return n_ - 1; }
    int n_markers() const { return int(vol_.size()); }
    int obs_size() const { return 8 + 2 * (n_ - 1); }

    void reset() {
        std::fill(q_.begin(), q_.end(), 0.0);
        std::fill(qd_.begin(), qd_.end(), 0.0);
        yaw_ = d_.heading; yawrate_ = 0; px_ = pz_ = 0; Ly_ = 0; dpx_ = dpz_ = dLy_ = 0;
        kinematics(yaw_);
        comx_ = d_.root_pos[0] + offx_; comz_ = d_.root_pos[2] + offz_;
        assemble();
    }

    // wrench6[k] = (force, torque about origin3[k]) the fluid exerted on link k during the previous substep
    void advance(const float *action, const double *wrench6, const double *origin3) {
        for (int j = 0; j < n_ - 1; ++j) {
            const double goal = double(action[j]) * d_.joint_limit;
            double rate = d_.joint_gain * (goal - q_[j]);
            rate = std::max(-d_.joint_rate_max, std::min(d_.joint_rate_max, rate));
            const double qn = std::max(-d_.joint_limit, std::min(d_.joint_limit, q_[j] + rate));
            qd_[j] = qn - q_[j];
            q_[j] = qn;
        }
        if (d_.free_root) {
            double Fx = 0, Fz = 0, Ty = 0;
            for (int k = 0; k < n_; ++k) {
                const double *w = wrench6 + 6 * k;
                const double ax = origin3[3 * k] - comx_, az = origin3[3 * k + 2] - comz_;
                Fx += w[0]; Fz += w[2];
                Ty += w[4] + (az * w[0] - ax * w[2]);      // (arm x F).y
            }
            kinematics(yaw_);                               // old heading, new joint angles
            double Itot = 0, Lshape = 0, Crot = 0;
            for (int k = 0; k < n_; ++k) {
                const double r2 = rx_[k] * rx_[k] + rz_[k] * rz_[k];
                Itot += I_[k] + m_[k] * r2;
                Lshape += I_[k] * wrel_[k] + m_[k] * (rz_[k] * sx_[k] - rx_[k] * sz_[k]);
                Crot += pen_[k] * r2 + pen2_[k];
            }
            // Added-mass stabilisation of the explicit body<->fluid coupling: the direct-forcing penalty
            // F = 2(U_d - U*) is stiff against the body velocity (stiffness sum 2 dV).  A virtual mass
            // Mv = beta * stiffness low-pass filters the momentum increment, (M+Mv) a_new = F + Mv a_old;
            // the fixed point is a = F/M and the update is unconditionally stable for beta >= 1/4.
            // With n direct-forcing passes per substep (FgConfig.ib_iterations) the virtual mass is 2^n - 1 times as large (set_forcing_passes).
            const double Mv = kBeta * passes_ * pen_total_, Iv = kBeta * passes_ * Crot;
            dpx_ = (M_ * Fx + Mv * dpx_) / (M_ + Mv);
            dpz_ = (M_ * Fz + Mv * dpz_) / (M_ + Mv);
            dLy_ = (Itot * Ty + Iv * dLy_) / (Itot + Iv);
            px_ += dpx_; pz_ += dpz_; Ly_ += dLy_;
            comx_ += px_ / M_; comz_ += pz_ / M_;
            yawrate_ = (Ly_ - Lshape) / Itot;
            yaw_ += yawrate_;
        }
        kinematics(yaw_);
        assemble();
    }

    void emit_markers(float *X, float *U, float *dV, int32_t *link, int link_offset, double *origin3) const {
        for (int k = 0; k < n_; ++k) {
            const double tx = std::sin(ang_[k]), tz = std::cos(ang_[k]);   // tail-ward axis
            const double ex = tz, ez = -tx;                                // lateral axis
            origin3[3 * k] = cx_[k]; origin3[3 * k + 1] = d_.root_pos[1]; origin3[3 * k + 2] = cz_[k];
            for (int i = start_[k]; i < start_[k + 1]; ++i) {
                const double a = pts_[3 * i], b = pts_[3 * i + 1], c = pts_[3 * i + 2];
                const double rx = a * ex + c * tx, rz = a * ez + c * tz;
                X[3 * i] = float(cx_[k] + rx);
                X[3 * i + 1] = float(d_.root_pos[1] + b);
                X[3 * i + 2] = float(cz_[k] + rz);
                U[3 * i] = float(vx_[k] + w_[k] * rz);      // omega y_hat x r = omega (r_z, 0, -r_x)
                U[3 * i + 1] = 0.f;
                U[3 * i + 2] = float(vz_[k] + w_[k] * (-rx));
                dV[i] = float(vol_[i]);
                link[i] = link_offset + k;
            }
        }
    }

    void write_obs(float *o) const {
        const int nj = n_ - 1;
        o[0] = float(cx_[0]); o[1] = float(d_.root_pos[1]); o[2] = float(cz_[0]);
        o[3] = float(yaw_);
        o[4] = float(px_ / M_); o[5] = 0.f; o[6] = float(pz_ / M_);
        o[7] = float(yawrate_);
        for (int j = 0; j < nj; ++j) { o[8 + j] = float(q_[j]); o[8 + nj + j] = float(qd_[j]); }
    }

private:
    // chain geometry for a given root heading: link centres (r) and shape-change velocities (s) relative to the
    // centre of mass, relative yaw rates, absolute link angles
    void kinematics(double yaw) {
        double ang = yaw, wr = 0, x = 0, z = 0, ux = 0, uz = 0;
        double mx = 0, mz = 0, mux = 0, muz = 0;
        for (int k = 0; k < n_; ++k) {
            if (k > 0) {
                const double a0 = half_[k - 1], a1 = half_[k];
                double tx = std::sin(ang), tz = std::cos(ang);
                x += a0 * tx; z += a0 * tz;
                ux += wr * (a0 * tz); uz += wr * (-(a0 * tx));
                ang += q_[k - 1]; wr += qd_[k - 1];
                tx = std::sin(ang); tz = std::cos(ang);
                x += a1 * tx; z += a1 * tz;
                ux += wr * (a1 * tz); uz += wr * (-(a1 * tx));
            }
            ang_[k] = ang; wrel_[k] = wr; rx_[k] = x; rz_[k] = z; sx_[k] = ux; sz_[k] = uz;
            mx += m_[k] * x; mz += m_[k] * z; mux += m_[k] * ux; muz += m_[k] * uz;
        }
        offx_ = mx / M_; offz_ = mz / M_;
        for (int k = 0; k < n_; ++k) {
            rx_[k] -= offx_; rz_[k] -= offz_;
            sx_[k] -= mux / M_; sz_[k] -= muz / M_;
        }
    }
    void assemble() {
        double vcx = px_ / M_, vcz = pz_ / M_;
        if (!d_.free_root) {   // pinned: head centre fixed at root_pos
            comx_ = d_.root_pos[0] + offx_; comz_ = d_.root_pos[2] + offz_;
            vcx = -sx_[0]; vcz = -sz_[0];
        }
        for (int k = 0; k < n_; ++k) {
            cx_[k] = comx_ + rx_[k]; cz_[k] = comz_ + rz_[k];
            vx_[k] = vcx + yawrate_ * rz_[k] + sx_[k];
            vz_[k] = vcz + yawrate_ * (-rx_[k]) + sz_[k];
            w_[k] = yawrate_ + wrel_[k];
        }
    }

    FgFishDesc d_{};
    int n_ = 0;
    double M_ = 0;
    std::vector<double> half_, m_, I_;
    std::vector<int> start_;
    std::vector<double> pts_, vol_;
    std::vector<double> q_, qd_;
    static constexpr double kBeta = 0.5;
    double passes_ = 1.0;                  // direct-forcing passes per substep (set_forcing_passes)
    std::vector<double> pen_, pen2_;       // per link: sum 2 dV and sum 2 dV |xi|^2 (in the swimming plane)
    double pen_total_ = 0;
    double dpx_ = 0, dpz_ = 0, dLy_ = 0;   // filtered momentum increments of the last substep
    double yaw_ = 0, yawrate_ = 0, px_ = 0, pz_ = 0, Ly_ = 0, comx_ = 0, comz_ = 0, offx_ = 0, offz_ = 0;
    std::vector<double> rx_, rz_, sx_, sz_, wrel_, ang_, cx_, cz_, vx_, vz_, w_;
};

}  // namespace fg
