// Host-side URDF reader (see model.h).  Self-contained: a small XML element scanner is enough for URDF.
#include "model.h"

#include <strings.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace dwbc {
namespace {

struct Xml {
    std::string tag;
    std::map<std::string, std::string> attr;
    std::vector<std::unique_ptr<Xml>> kids;
    const Xml *child(const char *t) const {
        for (auto &k : kids)
            if (k->tag == t) return k.get();
        return nullptr;
    }
    const char *get(const char *a) const {
        auto it = attr.find(a);
        return it == attr.end() ? nullptr : it->second.c_str();
    }
};

struct XmlParser {
    const std::string &s;
    size_t i = 0;
    explicit XmlParser(const std::string &src) : s(src) {}
    void skip_ws() {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
    }
    bool starts(const char *p) const { return s.compare(i, strlen(p), p) == 0; }
    void skip_misc() {  // whitespace, text, comments, declarations
        while (i < s.size()) {
            if (s[i] != '<') {
                i++;
            } else if (starts("<!--")) {
                size_t e = s.find("-->", i);
                if (e == std::string::npos) throw std::runtime_error("urdf: unterminated comment");
                i = e + 3;
            } else if (starts("<?") || starts("<!")) {
                size_t e = s.find('>', i);
                if (e == std::string::npos) throw std::runtime_error("urdf: unterminated declaration");
                i = e + 1;
            } else
                return;
        }
    }
    std::string name() {
        size_t b = i;
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == ':' || s[i] == '-' || s[i] == '.')) i++;
        return s.substr(b, i - b);
    }
    std::unique_ptr<Xml> element() {  // at '<' of an opening tag
        std::unique_ptr<Xml> e(new Xml());
        i++;
        e->tag = name();
        while (true) {
            skip_ws();
            if (i >= s.size()) throw std::runtime_error("urdf: unexpected end of file");
            if (s[i] == '/') {
                i += 2;
                return e;
            }
            if (s[i] == '>') {
                i++;
                break;
            }
            std::string a = name();
            skip_ws();
            if (a.empty() || s[i] != '=') throw std::runtime_error("urdf: malformed attribute in <" + e->tag + ">");
            i++;
            skip_ws();
            char q = s[i++];
            size_t b = i;
            while (i < s.size() && s[i] != q) i++;
            e->attr[a] = s.substr(b, i - b);
            i++;
        }
        while (true) {
            skip_misc();
            if (i >= s.size()) throw std::runtime_error("urdf: missing </" + e->tag + ">");
            if (starts("</")) {
                size_t c = s.find('>', i);
                i = c + 1;
                return e;
            }
            e->kids.push_back(element());
        }
    }
};

struct V3 {
    double v[3] = {0, 0, 0};
};
struct M3 {
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
};
M3 mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
M3 tr(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[j * 3 + i];
    return r;
}
V3 mul(const M3 &a, const V3 &x) {
    V3 r;
    for (int i = 0; i < 3; i++) r.v[i] = a.m[i * 3] * x.v[0] + a.m[i * 3 + 1] * x.v[1] + a.m[i * 3 + 2] * x.v[2];
    return r;
}
V3 parse3(const char *txt) {
    V3 r;
    if (txt) {
        std::istringstream is(txt);
        is >> r.v[0] >> r.v[1] >> r.v[2];
    }
    return r;
}
// URDF rpy: R = Rz(yaw) Ry(pitch) Rx(roll)
M3 rpy_rot(const V3 &rpy) {
    double cr = std::cos(rpy.v[0]), sr = std::sin(rpy.v[0]), cp = std::cos(rpy.v[1]), sp = std::sin(rpy.v[1]);
    double cy = std::cos(rpy.v[2]), sy = std::sin(rpy.v[2]);
    M3 Rx, Ry, Rz;
    Rx.m[4] = cr; Rx.m[5] = -sr; Rx.m[7] = sr; Rx.m[8] = cr;
    Ry.m[0] = cp; Ry.m[2] = sp;  Ry.m[6] = -sp; Ry.m[8] = cp;
    Rz.m[0] = cy; Rz.m[1] = -sy; Rz.m[3] = sy; Rz.m[4] = cy;
    return mul(Rz, mul(Ry, Rx));
}

struct LinkIn {
    double mass = 0;
    V3 com;
    M3 inertia;  // about COM, link frame
    bool has_inertia = false;
};
struct JointIn {
    std::string name, type, parent, child;
    V3 xyz, axis;
    M3 R;
};
struct BodyAcc {
    std::string name;
    int parent;
    M3 E;
    V3 r, axis;
    double mass;
    V3 com;
    M3 inertia;
};

// rigidly attach `l` at pose (R, r) in the body frame (fixed joints, merged like RBDL does)
void merge_link(BodyAcc &b, const M3 &R, const V3 &r, const LinkIn &l) {
    if (l.mass <= 0) return;
    V3 c2 = mul(R, l.com);
    for (int k = 0; k < 3; k++) c2.v[k] += r.v[k];
    M3 I2 = mul(R, mul(l.inertia, tr(R)));
    double m1 = b.mass, m2 = l.mass, mt = m1 + m2;
    V3 c;
    for (int k = 0; k < 3; k++) c.v[k] = (m1 * b.com.v[k] + m2 * c2.v[k]) / mt;
    auto shift = [&](const M3 &I, double m, const V3 &from) {
        V3 d;
        for (int k = 0; k < 3; k++) d.v[k] = from.v[k] - c.v[k];
        double dd = d.v[0] * d.v[0] + d.v[1] * d.v[1] + d.v[2] * d.v[2];
        M3 out = I;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) out.m[i * 3 + j] += m * ((i == j ? dd : 0.0) - d.v[i] * d.v[j]);
        return out;
    };
    M3 a = shift(b.inertia, m1, b.com), bb = shift(I2, m2, c2);
    for (int k = 0; k < 9; k++) b.inertia.m[k] = a.m[k] + bb.m[k];
    b.mass = mt;
    b.com = c;
}

}  // namespace

int HostModel::link_id(const char *name) const {
    if (!name) return -1;
    if (strcasecmp(name, "COM") == 0) return dev.nb;
    for (size_t i = 0; i < names.size(); i++)
        if (strcasecmp(names[i].c_str(), name) == 0) return (int)i;
    return -1;
}

HostModel load_urdf_model(const std::string &path, bool floating) {
    if (!floating) throw std::runtime_error("only floating-base models are supported (the reference hot path always loads with floating=true)");
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open urdf: " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    std::string src = ss.str();
    XmlParser p(src);
    p.skip_misc();
    if (p.i >= src.size()) throw std::runtime_error("urdf: no root element");
    std::unique_ptr<Xml> root = p.element();
    if (root->tag != "robot") throw std::runtime_error("urdf: root element is not <robot>");

    std::map<std::string, LinkIn> links;
    std::vector<std::string> link_order;
    std::map<std::string, std::vector<JointIn>> kids;
    std::map<std::string, bool> is_child;
    for (auto &e : root->kids) {
        if (e->tag == "link") {
            LinkIn l;
            const Xml *in_ = e->child("inertial");
            if (in_) {
                const Xml *o = in_->child("origin");
                M3 Ri;
                if (o) {
                    l.com = parse3(o->get("xyz"));
                    Ri = rpy_rot(parse3(o->get("rpy")));
                }
                const Xml *ms = in_->child("mass");
                if (ms && ms->get("value")) l.mass = atof(ms->get("value"));
                const Xml *it = in_->child("inertia");
                if (it) {
                    auto g = [&](const char *k) { return it->get(k) ? atof(it->get(k)) : 0.0; };
                    M3 I;
                    I.m[0] = g("ixx"); I.m[1] = g("ixy"); I.m[2] = g("ixz");
                    I.m[3] = g("ixy"); I.m[4] = g("iyy"); I.m[5] = g("iyz");
                    I.m[6] = g("ixz"); I.m[7] = g("iyz"); I.m[8] = g("izz");
                    l.inertia = mul(Ri, mul(I, tr(Ri)));
                    l.has_inertia = true;
                }
            }
            if (!l.has_inertia)
                for (int k = 0; k < 9; k++) l.inertia.m[k] = 0;
            const char *nm = e->get("name");
            if (!nm) throw std::runtime_error("urdf: <link> without name");
            links[nm] = l;
            link_order.push_back(nm);
        } else if (e->tag == "joint") {
            JointIn j;
            j.name = e->get("name") ? e->get("name") : "";
            j.type = e->get("type") ? e->get("type") : "";
            const Xml *pa = e->child("parent"), *ch = e->child("child");
            if (!pa || !ch || !pa->get("link") || !ch->get("link")) throw std::runtime_error("urdf: joint " + j.name + " lacks parent/child");
            j.parent = pa->get("link");
            j.child = ch->get("link");
            const Xml *o = e->child("origin");
            if (o) {
                j.xyz = parse3(o->get("xyz"));
                j.R = rpy_rot(parse3(o->get("rpy")));
            }
            const Xml *ax = e->child("axis");
            j.axis.v[0] = 1;
            if (ax && ax->get("xyz")) j.axis = parse3(ax->get("xyz"));
            kids[j.parent].push_back(j);
            is_child[j.child] = true;
        }
    }
    std::string rootlink;
    for (auto &n : link_order)
        if (!is_child.count(n)) {
            if (!rootlink.empty()) throw std::runtime_error("urdf: more than one root link");
            rootlink = n;
        }
    if (rootlink.empty()) throw std::runtime_error("urdf: no root link");
    for (auto &kv : kids) std::sort(kv.second.begin(), kv.second.end(), [](const JointIn &a, const JointIn &b) { return a.name < b.name; });

    std::vector<BodyAcc> bodies;
    {
        BodyAcc b0;
        b0.name = rootlink;
        b0.parent = -1;
        b0.mass = links[rootlink].mass;
        b0.com = links[rootlink].com;
        b0.inertia = links[rootlink].inertia;
        bodies.push_back(b0);
    }
    // depth-first, iterative with an explicit stack to keep visiting order = recursive order
    struct Frame {
        std::string link;
        int body;
        M3 R;
        V3 r;
        size_t next;
    };
    std::vector<Frame> stack;
    stack.push_back({rootlink, 0, M3(), V3(), 0});
    while (!stack.empty()) {
        Frame &f = stack.back();
        auto it = kids.find(f.link);
        if (it == kids.end() || f.next >= it->second.size()) {
            stack.pop_back();
            continue;
        }
        const JointIn &j = it->second[f.next++];
        M3 Rj = mul(f.R, j.R);
        V3 rj = mul(f.R, j.xyz);
        for (int k = 0; k < 3; k++) rj.v[k] += f.r.v[k];
        if (!links.count(j.child)) throw std::runtime_error("urdf: joint " + j.name + " references unknown link " + j.child);
        if (j.type == "fixed") {
            int body = f.body;
            merge_link(bodies[body], Rj, rj, links[j.child]);
            stack.push_back({j.child, body, Rj, rj, 0});
        } else if (j.type == "revolute" || j.type == "continuous") {
            BodyAcc b;
            b.name = j.child;
            b.parent = f.body;
            b.E = Rj;
            b.r = rj;
            double an = std::sqrt(j.axis.v[0] * j.axis.v[0] + j.axis.v[1] * j.axis.v[1] + j.axis.v[2] * j.axis.v[2]);
            if (an == 0) throw std::runtime_error("urdf: joint " + j.name + " has a zero axis");
            for (int k = 0; k < 3; k++) b.axis.v[k] = j.axis.v[k] / an;
            b.mass = links[j.child].mass;
            b.com = links[j.child].com;
            b.inertia = links[j.child].inertia;
            bodies.push_back(b);
            stack.push_back({j.child, (int)bodies.size() - 1, M3(), V3(), 0});
        } else {
            throw std::runtime_error("urdf: unsupported joint type '" + j.type + "' on " + j.name);
        }
    }
    if ((int)bodies.size() > MAXB) throw std::runtime_error("urdf: too many bodies for this build (MAXB)");

    HostModel hm;
    DevModel &d = hm.dev;
    std::memset(&d, 0, sizeof(d));
    d.nb = (int)bodies.size();
    d.n = d.nb + 5;
    d.m = d.nb - 1;
    for (int i = 0; i < d.nb; i++) {
        const BodyAcc &b = bodies[i];
        hm.names.push_back(b.name);
        d.parent[i] = b.parent;
        d.depth[i] = b.parent < 0 ? 0 : d.depth[b.parent] + 1;
        d.max_depth = std::max(d.max_depth, d.depth[i]);
        d.anc_mask[i] = (b.parent < 0 ? 0ull : d.anc_mask[b.parent]) | (i > 0 ? (1ull << i) : 0ull);
        std::memcpy(d.E_T[i], b.E.m, sizeof(double) * 9);
        std::memcpy(d.r_T[i], b.r.v, sizeof(double) * 3);
        std::memcpy(d.axis[i], b.axis.v, sizeof(double) * 3);
        d.mass[i] = b.mass;
        std::memcpy(d.com[i], b.com.v, sizeof(double) * 3);
        std::memcpy(d.inertia[i], b.inertia.m, sizeof(double) * 9);
        d.total_mass += b.mass;
    }
    if (d.max_depth >= MAX_DEPTH) throw std::runtime_error("urdf: kinematic tree too deep for this build (MAX_DEPTH)");
    int pos = 0;
    for (int dep = 0; dep <= d.max_depth; dep++) {
        d.level_start[dep] = pos;
        for (int i = 0; i < d.nb; i++)
            if (d.depth[i] == dep) d.order[pos++] = i;
    }
    d.level_start[d.max_depth + 1] = pos;
    pos = 0;
    for (int b = 0; b < d.nb; b++) {
        d.child_start[b] = pos;
        for (int i = 0; i < d.nb; i++)
            if (d.parent[i] == b) d.child_list[pos++] = i;
    }
    d.child_start[d.nb] = pos;
    return hm;
}

}  // namespace dwbc
