// meshtools — host-side input generator for the icsb200 hot path (NOT part of the timed path).
//
// Builds OpenFOAM-style polyMesh connectivity + fvMesh geometry for block-structured hex meshes
// (the "synthetic blockMesh refinements of the shipped tutorials" of BASELINE.json) and for
// point/face lists read from a polyMesh directory.  Geometry follows OpenFOAM-v2112
// primitiveMesh/surfaceInterpolation semantics as restated in SURVEY.md Appendix A
// (face centre/area by triangle fan about the point average, cell centre/volume by face pyramids,
// linear weights, deltaCoeffs, nonOrthDeltaCoeffs).  In a real drop-in these arrays come straight
// from fvMesh (mesh.Sf(), mesh.magSf(), mesh.weights() ... see INTEGRATION.md); this file only
// exists because OpenFOAM is not available in the build image.
//
// Reference call sites that consume these arrays: hllcFluxScheme.C:78 (Sf/magSf),
// convectiveFluxScheme.C:65,376 (weights), setCoAndDeltaT.H:61 (nonOrthDeltaCoeffs),
// residualsUpdate.H:81-83 (V), lusgs.C:141-156 (owner/neighbour, upper-triangular order).
//
// C API (ctypes): icsmesh_* below.  All arrays are AoS in OpenFOAM's native layout
// (vector = 3 contiguous doubles).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double mag(V3 a) { return std::sqrt(dot(a, a)); }

constexpr double VSMALL = 1e-300, ROOTVSMALL = 1e-150;

enum PatchKind { PK_PATCH = 0, PK_WALL = 1, PK_EMPTY = 2, PK_SYMMETRYPLANE = 3, PK_CYCLIC = 4, PK_PROCESSOR = 5 };

struct Patch {
    std::string name;
    int kind = PK_PATCH;
    int start = 0, size = 0;
    int nbrRank = -1;   // processor patches
    int nbrPatch = -1;  // cyclic
};

struct Mesh {
    int nCells = 0, nInternalFaces = 0, nFaces = 0;
    std::vector<int> owner, neighbour;
    std::vector<V3> Sf, Cf, C;
    std::vector<double> magSf, V, weights, deltaCoeffs, nonOrthDeltaCoeffs;
    std::vector<Patch> patches;
    int solutionD[3] = {1, 1, 1};
    // processor-patch bookkeeping for partitions (global ids, ordered identically on both sides)
    std::vector<int> cellGlobal;        // local cell -> global cell
    std::vector<int> faceGlobal;        // local face -> global face (parent mesh)
    std::string error;
};

// ---------------------------------------------------------------- geometry
// face centre & area vector (SURVEY Appendix A "Face centre / area vector")
void faceGeom(const V3* p, int n, V3& Cf, V3& Sf)
{
    if (n == 3) {
        Cf = (1.0 / 3.0) * (p[0] + p[1] + p[2]);
        Sf = 0.5 * cross(p[1] - p[0], p[2] - p[0]);
        return;
    }
    V3 sumN{0, 0, 0}, sumAc{0, 0, 0}, fC{0, 0, 0};
    double sumA = 0;
    for (int i = 0; i < n; i++) fC = fC + p[i];
    fC = (1.0 / n) * fC;
    for (int i = 0; i < n; i++) {
        const V3& a = p[i];
        const V3& b = p[(i + 1) % n];
        V3 c = a + b + fC;
        V3 nn = cross(b - a, fC - a);
        double ar = mag(nn);
        sumN = sumN + nn;
        sumA += ar;
        sumAc = sumAc + ar * c;
    }
    if (sumA < ROOTVSMALL) { Cf = fC; Sf = {0, 0, 0}; }
    else { Cf = (1.0 / 3.0) * ((1.0 / sumA) * sumAc); Sf = 0.5 * sumN; }
}

// cell centres/volumes + interpolation coefficients from face geometry
void finishGeometry(Mesh& m)
{
    const int N = m.nCells, F = m.nInternalFaces, FT = m.nFaces;
    m.magSf.resize(FT);
    for (int f = 0; f < FT; f++) m.magSf[f] = mag(m.Sf[f]);
    std::vector<V3> cEst(N, V3{0, 0, 0});
    std::vector<int> nCellFaces(N, 0);
    for (int f = 0; f < FT; f++) { cEst[m.owner[f]] = cEst[m.owner[f]] + m.Cf[f]; nCellFaces[m.owner[f]]++; }
    for (int f = 0; f < F; f++) { cEst[m.neighbour[f]] = cEst[m.neighbour[f]] + m.Cf[f]; nCellFaces[m.neighbour[f]]++; }
    for (int c = 0; c < N; c++) cEst[c] = (1.0 / nCellFaces[c]) * cEst[c];
    m.C.assign(N, V3{0, 0, 0});
    m.V.assign(N, 0.0);
    for (int f = 0; f < FT; f++) {
        int o = m.owner[f];
        double pyr3 = dot(m.Sf[f], m.Cf[f] - cEst[o]);
        V3 pc = 0.75 * m.Cf[f] + 0.25 * cEst[o];
        m.C[o] = m.C[o] + pyr3 * pc;
        m.V[o] += pyr3;
    }
    for (int f = 0; f < F; f++) {
        int n = m.neighbour[f];
        double pyr3 = dot(m.Sf[f], cEst[n] - m.Cf[f]);
        V3 pc = 0.75 * m.Cf[f] + 0.25 * cEst[n];
        m.C[n] = m.C[n] + pyr3 * pc;
        m.V[n] += pyr3;
    }
    for (int c = 0; c < N; c++) {
        if (std::fabs(m.V[c]) > VSMALL) m.C[c] = (1.0 / m.V[c]) * m.C[c];
        else m.C[c] = cEst[c];
        m.V[c] *= (1.0 / 3.0);
    }
    m.weights.resize(FT);
    m.deltaCoeffs.resize(FT);
    m.nonOrthDeltaCoeffs.resize(FT);
    for (int f = 0; f < F; f++) {
        int o = m.owner[f], n = m.neighbour[f];
        double SfdOwn = std::fabs(dot(m.Sf[f], m.Cf[f] - m.C[o]));
        double SfdNei = std::fabs(dot(m.Sf[f], m.C[n] - m.Cf[f]));
        m.weights[f] = SfdNei / (SfdOwn + SfdNei);
        V3 d = m.C[n] - m.C[o];
        m.deltaCoeffs[f] = 1.0 / mag(d);
        V3 nh = (1.0 / m.magSf[f]) * m.Sf[f];
        m.nonOrthDeltaCoeffs[f] = 1.0 / std::max(dot(nh, d), 0.05 * mag(d));
    }
    for (int f = F; f < FT; f++) {
        int o = m.owner[f];
        m.weights[f] = 1.0;
        // fvPatch::delta() of a non-coupled patch is the PATCH-NORMAL delta nHat (nHat & (Cf - Cn)) (OpenFOAM fvPatch.C), so both
        // coefficients are 1 / (nHat & (Cf - Cn)); coupled patches overwrite theirs with the cell-to-cell delta (cyclicGeometry, extract_part)
        V3 d0 = m.Cf[f] - m.C[o];
        if (mag(d0) < VSMALL || m.magSf[f] < VSMALL) { m.deltaCoeffs[f] = 0; m.nonOrthDeltaCoeffs[f] = 0; continue; }
        V3 nh = (1.0 / m.magSf[f]) * m.Sf[f];
        V3 d = dot(nh, d0) * nh;
        double md = mag(d);
        if (md < VSMALL) { m.deltaCoeffs[f] = 0; m.nonOrthDeltaCoeffs[f] = 0; continue; }
        m.deltaCoeffs[f] = 1.0 / md;
        m.nonOrthDeltaCoeffs[f] = 1.0 / std::max(dot(nh, d), 0.05 * md);
    }
}

// weights / deltaCoeffs / nonOrthDeltaCoeffs of a translational cyclic pair (face i of patch ia matches face i of patch ib):
// cyclicFvPatch::makeWeights (own / neighbour normal distances) and cyclicFvPatch::delta() = own delta - neighbour delta
void cyclicGeometry(Mesh& m, int ia, int ib)
{
    for (int side = 0; side < 2; side++) {
        const Patch& pa = m.patches[side == 0 ? ia : ib];
        const Patch& pb = m.patches[side == 0 ? ib : ia];
        for (int i = 0; i < pa.size; i++) {
            const int fa = pa.start + i, fb = pb.start + i;
            const V3 nfa = (1.0 / m.magSf[fa]) * m.Sf[fa], nfb = (1.0 / m.magSf[fb]) * m.Sf[fb];
            const V3 dA = m.Cf[fa] - m.C[m.owner[fa]], dB = m.Cf[fb] - m.C[m.owner[fb]];
            const double da = std::fabs(dot(nfa, dA)), db = std::fabs(dot(nfb, dB));
            m.weights[fa] = db / (da + db);
            const V3 d = dA - dB;
            const double md = mag(d);
            m.deltaCoeffs[fa] = 1.0 / md;
            m.nonOrthDeltaCoeffs[fa] = 1.0 / std::max(dot(nfa, d), 0.05 * md);
        }
    }
}

// ---------------------------------------------------------------- structured generator
// A block-structured hex mesh made of nb blocks stacked along x, each (nxb x ny x nz) cells,
// numbered block by block with i fastest (blockMesh cell numbering).  Points come from a mapping
// of the logical coordinates (xi, eta, zeta) in [0,1]^3.
struct Mapping {
    int kind;  // 0 box, 1 circular-arc bump channel, 2 box with sinusoidal bump on z-min
    double lo[3], hi[3];
    double gradY;  // simpleGrading expansion ratio in y (1 = uniform)
    double amp;    // bump height (kind 2)
};

inline double gradedLambda(int j, int n, double r)
{
    if (r == 1.0 || n <= 1) return double(j) / n;
    double k = std::pow(r, 1.0 / (n - 1));
    return (1.0 - std::pow(k, j)) / (1.0 - std::pow(k, n));
}

struct StructGen {
    int nb, nxb, ny, nz;
    Mapping map;
    std::vector<V3> pts;  // cached points, (nxT+1)*(ny+1)*(nz+1)
    // sub-box of a larger global grid: this generator covers global cells [off, off + n) in each direction
    int off[3] = {0, 0, 0};
    int glob[3] = {0, 0, 0};  // global cell counts (0 = same as local)
    int nxT() const { return nb * nxb; }
    int gx() const { return glob[0] ? glob[0] : nxT(); }
    int gy() const { return glob[1] ? glob[1] : ny; }
    int gz() const { return glob[2] ? glob[2] : nz; }
    void cachePoints()
    {
        const int nx = nxT();
        pts.resize((size_t)(nx + 1) * (ny + 1) * (nz + 1));
        for (int K = 0; K <= nz; K++)
            for (int J = 0; J <= ny; J++)
                for (int I = 0; I <= nx; I++) pts[((size_t)K * (ny + 1) + J) * (nx + 1) + I] = computePoint(I, J, K);
    }
    V3 point(int I, int J, int K) const { return pts[((size_t)K * (ny + 1) + J) * (nxT() + 1) + I]; }
    V3 computePoint(int Il, int Jl, int Kl) const
    {
        const int I = Il + off[0], J = Jl + off[1], K = Kl + off[2];
        double xi = double(I) / gx();
        double eta = gradedLambda(J, gy(), map.gradY);
        double zeta = double(K) / gz();
        V3 p;
        p.x = map.lo[0] + xi * (map.hi[0] - map.lo[0]);
        p.y = map.lo[1] + eta * (map.hi[1] - map.lo[1]);
        p.z = map.lo[2] + zeta * (map.hi[2] - map.lo[2]);
        if (map.kind == 1) {
            // circularArcBump: bottom wall is a circular arc of height amp (=0.1) over the middle
            // third of the channel (tutorials/circularArcBump/transonic/system/blockMeshDict:44-53).
            double L = map.hi[0] - map.lo[0];
            double xa = map.lo[0] + L / 3.0, xb = map.lo[0] + 2.0 * L / 3.0;
            double yb = map.lo[1];
            double xbot = p.x;
            if (p.x > xa && p.x < xb) {
                // arc through (xa,0),(mid,amp),(xb,0): radius R, centre (mid, amp-R); uniform in angle
                double half = 0.5 * (xb - xa), h = map.amp;
                double R = (half * half + h * h) / (2.0 * h);
                double th0 = std::asin(half / R);
                double s = (p.x - xa) / (xb - xa);  // block-local xi
                double th = -th0 + 2.0 * th0 * s;
                xbot = 0.5 * (xa + xb) + R * std::sin(th);
                yb = map.lo[1] + (h - R) + R * std::cos(th);
            }
            // linear blend bottom edge point -> top edge point (straight top wall)
            p.x = xbot + eta * (p.x - xbot);
            p.y = yb + eta * (map.hi[1] - yb);
        } else if (map.kind == 2) {
            // synthetic "wing-like" disturbance: smooth bump on the z-min wall decaying to z-max
            double sx = (p.x - map.lo[0]) / (map.hi[0] - map.lo[0]);
            double sy = (p.y - map.lo[1]) / (map.hi[1] - map.lo[1]);
            double bx = (sx > 0.3 && sx < 0.7) ? std::pow(std::sin(M_PI * (sx - 0.3) / 0.4), 2) : 0.0;
            double by = (sy < 0.6) ? std::pow(std::cos(0.5 * M_PI * sy / 0.6), 2) : 0.0;
            p.z += map.amp * bx * by * (1.0 - zeta);
        }
        return p;
    }
    int cellId(int I, int j, int k) const
    {
        int b = I / nxb, i = I % nxb;
        return b * (nxb * ny * nz) + i + nxb * (j + ny * k);
    }
};

// patch order and orientation for the six sides: xmin,xmax,ymin,ymax,zmin,zmax
Mesh* buildStructured(const StructGen& g, const int kinds[6], const char* const names[6])
{
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    const int nx = g.nxT(), ny = g.ny, nz = g.nz;
    const long long N = 1LL * nx * ny * nz;
    const long long FI = 1LL * (nx - 1) * ny * nz + 1LL * nx * (ny - 1) * nz + 1LL * nx * ny * (nz - 1);
    const long long FB = 2LL * (ny * nz + nx * nz + nx * ny);
    if (N > 2000000000LL || FI + FB > 2000000000LL) { m.error = "mesh too large for 32-bit labels"; return mp; }
    m.nCells = (int)N;
    m.nInternalFaces = (int)FI;
    m.nFaces = (int)(FI + FB);
    m.owner.resize(m.nFaces);
    m.neighbour.resize(m.nInternalFaces);
    m.Sf.resize(m.nFaces);
    m.Cf.resize(m.nFaces);

    // inverse map cell id -> (I,j,k)
    auto ijk = [&](int c, int& I, int& j, int& k) {
        int per = g.nxb * ny * nz;
        int b = c / per, r = c % per;
        int i = r % g.nxb;
        j = (r / g.nxb) % ny;
        k = r / (g.nxb * ny);
        I = b * g.nxb + i;
    };
    // face quad of cell (I,j,k) in direction dir (0:+x 1:+y 2:+z 3:-x 4:-y 5:-z), outward normal
    auto faceQuad = [&](int I, int j, int k, int dir, V3 q[4]) {
        switch (dir) {
            case 0: q[0] = g.point(I + 1, j, k); q[1] = g.point(I + 1, j + 1, k); q[2] = g.point(I + 1, j + 1, k + 1); q[3] = g.point(I + 1, j, k + 1); break;
            case 3: q[0] = g.point(I, j, k); q[1] = g.point(I, j, k + 1); q[2] = g.point(I, j + 1, k + 1); q[3] = g.point(I, j + 1, k); break;
            case 1: q[0] = g.point(I, j + 1, k); q[1] = g.point(I, j + 1, k + 1); q[2] = g.point(I + 1, j + 1, k + 1); q[3] = g.point(I + 1, j + 1, k); break;
            case 4: q[0] = g.point(I, j, k); q[1] = g.point(I + 1, j, k); q[2] = g.point(I + 1, j, k + 1); q[3] = g.point(I, j, k + 1); break;
            case 2: q[0] = g.point(I, j, k + 1); q[1] = g.point(I + 1, j, k + 1); q[2] = g.point(I + 1, j + 1, k + 1); q[3] = g.point(I, j + 1, k + 1); break;
            default: q[0] = g.point(I, j, k); q[1] = g.point(I, j + 1, k); q[2] = g.point(I + 1, j + 1, k); q[3] = g.point(I + 1, j, k); break;
        }
    };
    // internal faces in upper-triangular order: for each cell ascending, higher neighbours ascending
    long long f = 0;
    for (int c = 0; c < m.nCells; c++) {
        int I, j, k;
        ijk(c, I, j, k);
        std::array<std::pair<int, int>, 3> nb;
        int n = 0;
        if (I + 1 < nx) nb[n++] = {g.cellId(I + 1, j, k), 0};
        if (j + 1 < ny) nb[n++] = {g.cellId(I, j + 1, k), 1};
        if (k + 1 < nz) nb[n++] = {g.cellId(I, j, k + 1), 2};
        std::sort(nb.begin(), nb.begin() + n);
        for (int t = 0; t < n; t++) {
            // all +dir neighbours have a higher id in this numbering except across block seams
            // where the +x neighbour id is also higher (next block) — assert that
            if (nb[t].first <= c) { m.error = "non upper-triangular numbering"; return mp; }
            V3 q[4];
            faceQuad(I, j, k, nb[t].second, q);
            m.owner[f] = c;
            m.neighbour[f] = nb[t].first;
            faceGeom(q, 4, m.Cf[f], m.Sf[f]);
            f++;
        }
    }
    // boundary patches
    for (int side = 0; side < 6; side++) {
        Patch p;
        p.name = names[side];
        p.kind = kinds[side];
        p.start = (int)f;
        int dir = (side == 0) ? 3 : (side == 1) ? 0 : (side == 2) ? 4 : (side == 3) ? 1 : (side == 4) ? 5 : 2;
        int n1 = (side < 2) ? ny : nx, n2 = (side < 2) ? nz : (side < 4 ? nz : ny);
        for (int b = 0; b < n2; b++)
            for (int a = 0; a < n1; a++) {
                int I, j, k;
                if (side < 2) { I = (side == 0) ? 0 : nx - 1; j = a; k = b; }
                else if (side < 4) { I = a; j = (side == 2) ? 0 : ny - 1; k = b; }
                else { I = a; j = b; k = (side == 4) ? 0 : nz - 1; }
                V3 q[4];
                faceQuad(I, j, k, dir, q);
                m.owner[f] = g.cellId(I, j, k);
                faceGeom(q, 4, m.Cf[f], m.Sf[f]);
                f++;
            }
        p.size = (int)f - p.start;
        m.patches.push_back(p);
    }
    // faceCells of a patch must be usable as given; OpenFOAM does not require an order
    for (int d = 0; d < 3; d++) m.solutionD[d] = 1;
    if (kinds[0] == PK_EMPTY) m.solutionD[0] = -1;
    if (kinds[2] == PK_EMPTY) m.solutionD[1] = -1;
    if (kinds[4] == PK_EMPTY) m.solutionD[2] = -1;
    finishGeometry(m);
    return mp;
}

// ---------------------------------------------------------------- polyMesh reader
std::string slurp(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return {};
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

// strip C/C++ comments and the FoamFile header; return the text after the header
std::string stripFoam(const std::string& s)
{
    std::string o;
    o.reserve(s.size());
    for (size_t i = 0; i < s.size();) {
        if (s[i] == '/' && i + 1 < s.size() && s[i + 1] == '/') { while (i < s.size() && s[i] != '\n') i++; }
        else if (s[i] == '/' && i + 1 < s.size() && s[i + 1] == '*') { i += 2; while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) i++; i += 2; }
        else o.push_back(s[i++]);
    }
    size_t h = o.find("FoamFile");
    if (h != std::string::npos) { size_t e = o.find('}', h); if (e != std::string::npos) o = o.substr(e + 1); }
    return o;
}

Mesh* readPolyMesh(const std::string& dir)
{
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    std::string sp = stripFoam(slurp(dir + "/points")), sf = stripFoam(slurp(dir + "/faces")),
                so = stripFoam(slurp(dir + "/owner")), sn = stripFoam(slurp(dir + "/neighbour")),
                sb = stripFoam(slurp(dir + "/boundary"));
    if (sp.empty() || sf.empty() || so.empty() || sn.empty() || sb.empty()) { m.error = "cannot read polyMesh files in " + dir; return mp; }
    for (char& ch : sp) if (ch == '(' || ch == ')') ch = ' ';
    std::vector<V3> pts;
    { std::istringstream is(sp); long n; is >> n; pts.resize(n); for (long i = 0; i < n; i++) is >> pts[i].x >> pts[i].y >> pts[i].z; }
    std::vector<std::vector<int>> faces;
    {
        for (char& ch : sf) if (ch == '(' || ch == ')') ch = ' ';
        std::istringstream is(sf);
        long n; is >> n; faces.resize(n);
        for (long i = 0; i < n; i++) { int k; is >> k; faces[i].resize(k); for (int t = 0; t < k; t++) is >> faces[i][t]; }
    }
    auto readLabels = [](std::string s, std::vector<int>& v) {
        for (char& ch : s) if (ch == '(' || ch == ')') ch = ' ';
        std::istringstream is(s); long n; is >> n; v.resize(n); for (long i = 0; i < n; i++) is >> v[i];
    };
    readLabels(so, m.owner);
    readLabels(sn, m.neighbour);
    m.nFaces = (int)m.owner.size();
    m.nInternalFaces = (int)m.neighbour.size();
    m.nCells = 0;
    for (int o : m.owner) m.nCells = std::max(m.nCells, o + 1);
    // boundary
    {
        std::istringstream is(sb);
        int np; is >> np; std::string tok; is >> tok;  // "("
        std::vector<std::string> nbrNames;
        for (int p = 0; p < np; p++) {
            Patch pa; is >> pa.name; is >> tok;  // "{"
            std::string type, nbrName;
            int depth = 1;
            while (depth > 0 && (is >> tok)) {
                if (tok == "{") depth++;
                else if (tok == "}") depth--;
                else if (tok == "type") { is >> type; if (!type.empty() && type.back() == ';') type.pop_back(); }
                else if (tok == "neighbourPatch") { is >> nbrName; if (!nbrName.empty() && nbrName.back() == ';') nbrName.pop_back(); }
                else if (tok == "transform") { is >> tok; if (!tok.empty() && tok.back() == ';') tok.pop_back(); if (tok == "rotational") m.error = "rotational cyclic patch " + pa.name + " is not supported"; }
                else if (tok == "neighbProcNo") { is >> tok; pa.nbrRank = std::atoi(tok.c_str()); }   // processorPolyPatch (decomposePar output)
                else if (tok == "nFaces") { is >> tok; pa.size = std::atoi(tok.c_str()); }
                else if (tok == "startFace") { is >> tok; pa.start = std::atoi(tok.c_str()); }
            }
            pa.kind = type == "wall" ? PK_WALL : type == "empty" ? PK_EMPTY : type == "symmetryPlane" ? PK_SYMMETRYPLANE
                      : type == "cyclic" ? PK_CYCLIC : type == "processor" ? PK_PROCESSOR : PK_PATCH;
            if (type == "processorCyclic") m.error = "processorCyclic patch " + pa.name + " is not supported (decompose with preservePatches)";
            m.patches.push_back(pa);
            nbrNames.push_back(nbrName);
        }
        // cyclic pairs: neighbourPatch names -> patch indices (face i of a patch matches face i of its neighbour patch)
        for (int p = 0; p < np; p++)
            if (m.patches[p].kind == PK_CYCLIC) {
                for (int q = 0; q < np; q++) if (m.patches[q].name == nbrNames[p]) m.patches[p].nbrPatch = q;
                if (m.patches[p].nbrPatch < 0) m.error = "cyclic patch " + m.patches[p].name + " has no neighbourPatch";
            }
    }
    m.Sf.resize(m.nFaces);
    m.Cf.resize(m.nFaces);
    std::vector<V3> q;
    for (int f = 0; f < m.nFaces; f++) {
        q.resize(faces[f].size());
        for (size_t t = 0; t < q.size(); t++) q[t] = pts[faces[f][t]];
        faceGeom(q.data(), (int)q.size(), m.Cf[f], m.Sf[f]);
    }
    // solutionD: a direction is empty if an empty patch has normals along it
    double emptyN[3] = {0, 0, 0};
    for (auto& p : m.patches)
        if (p.kind == PK_EMPTY)
            for (int f = p.start; f < p.start + p.size; f++) {
                double ms = mag(m.Sf[f]);
                if (ms > 0) { emptyN[0] += std::fabs(m.Sf[f].x) / ms; emptyN[1] += std::fabs(m.Sf[f].y) / ms; emptyN[2] += std::fabs(m.Sf[f].z) / ms; }
            }
    double tot = emptyN[0] + emptyN[1] + emptyN[2];
    for (int d = 0; d < 3; d++) m.solutionD[d] = (tot > 0 && emptyN[d] > 0.5 * tot / 1.5 && emptyN[d] / tot > 0.9) ? -1 : 1;
    finishGeometry(m);
    for (int p = 0; p < (int)m.patches.size(); p++)
        if (m.patches[p].kind == PK_CYCLIC && p < m.patches[p].nbrPatch) {
            if (m.patches[p].size != m.patches[m.patches[p].nbrPatch].size) { m.error = "cyclic pair " + m.patches[p].name + " has patches of different size"; return mp; }
            cyclicGeometry(m, p, m.patches[p].nbrPatch);
        }
    return mp;
}

// ---------------------------------------------------------------- partitioning (decomposePar stand-in)
// Extract sub-mesh of cells with part[c]==rank.  Cells keep their relative global order, internal
// faces keep their relative order (which preserves upper-triangular ordering), original boundary
// patches keep their order (possibly empty), then one processor patch per neighbour rank in
// ascending rank order with faces in ascending global face order — the same order on both sides,
// as decomposePar produces.
Mesh* extractPart(const Mesh& g, const int* part, int rank)
{
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    std::vector<int> g2l(g.nCells, -1);
    for (int c = 0; c < g.nCells; c++) if (part[c] == rank) { g2l[c] = m.nCells++; m.cellGlobal.push_back(c); }
    auto push = [&](int gf, int own, bool flip) {
        m.owner.push_back(own);
        m.Sf.push_back(flip ? -1.0 * g.Sf[gf] : g.Sf[gf]);
        m.Cf.push_back(g.Cf[gf]);
        m.faceGlobal.push_back(gf);
    };
    for (int f = 0; f < g.nInternalFaces; f++) {
        int o = g.owner[f], n = g.neighbour[f];
        if (part[o] == rank && part[n] == rank) { push(f, g2l[o], false); m.neighbour.push_back(g2l[n]); }
    }
    m.nInternalFaces = (int)m.neighbour.size();
    for (auto& p : g.patches) {
        Patch q = p;
        q.start = (int)m.owner.size();
        for (int f = p.start; f < p.start + p.size; f++) if (part[g.owner[f]] == rank) push(f, g2l[g.owner[f]], false);
        q.size = (int)m.owner.size() - q.start;
        m.patches.push_back(q);
    }
    std::map<int, std::vector<int>> procFaces;
    for (int f = 0; f < g.nInternalFaces; f++) {
        int o = g.owner[f], n = g.neighbour[f];
        if (part[o] == rank && part[n] != rank) procFaces[part[n]].push_back(f);
        else if (part[n] == rank && part[o] != rank) procFaces[part[o]].push_back(f);
    }
    for (auto& kv : procFaces) {
        Patch q;
        q.name = "procBoundary" + std::to_string(rank) + "to" + std::to_string(kv.first);
        q.kind = PK_PROCESSOR;
        q.nbrRank = kv.first;
        q.start = (int)m.owner.size();
        for (int f : kv.second) {
            bool ownSide = part[g.owner[f]] == rank;
            push(f, g2l[ownSide ? g.owner[f] : g.neighbour[f]], !ownSide);
        }
        q.size = (int)m.owner.size() - q.start;
        m.patches.push_back(q);
    }
    m.nFaces = (int)m.owner.size();
    for (int d = 0; d < 3; d++) m.solutionD[d] = g.solutionD[d];
    finishGeometry(m);
    return mp;
}

// ---------------------------------------------------------------- renumbering (renumberMesh stand-in)
// new cell id = perm[old id].  Faces whose new owner would exceed the new neighbour are flipped, internal faces are
// re-sorted into upper-triangular order, boundary patches keep their order.
Mesh* renumber(const Mesh& g, const int* perm)
{
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    m.nCells = g.nCells; m.nInternalFaces = g.nInternalFaces; m.nFaces = g.nFaces;
    struct IF { int o, n, f; bool flip; };
    std::vector<IF> ifs(g.nInternalFaces);
    for (int f = 0; f < g.nInternalFaces; f++) {
        int o = perm[g.owner[f]], n = perm[g.neighbour[f]];
        bool flip = o > n;
        if (flip) std::swap(o, n);
        ifs[f] = {o, n, f, flip};
    }
    std::sort(ifs.begin(), ifs.end(), [](const IF& a, const IF& b) { return a.o != b.o ? a.o < b.o : (a.n != b.n ? a.n < b.n : a.f < b.f); });
    m.owner.resize(g.nFaces); m.neighbour.resize(g.nInternalFaces); m.Sf.resize(g.nFaces); m.Cf.resize(g.nFaces);
    for (int k = 0; k < g.nInternalFaces; k++) {
        m.owner[k] = ifs[k].o; m.neighbour[k] = ifs[k].n;
        m.Sf[k] = ifs[k].flip ? -1.0 * g.Sf[ifs[k].f] : g.Sf[ifs[k].f];
        m.Cf[k] = g.Cf[ifs[k].f];
    }
    for (int f = g.nInternalFaces; f < g.nFaces; f++) { m.owner[f] = perm[g.owner[f]]; m.Sf[f] = g.Sf[f]; m.Cf[f] = g.Cf[f]; }
    m.patches = g.patches;
    for (int d = 0; d < 3; d++) m.solutionD[d] = g.solutionD[d];
    finishGeometry(m);
    return mp;
}

}  // namespace

// ================================================================ C API
extern "C" {

void* icsmesh_structured(int nb, int nxb, int ny, int nz, int kind, const double lo[3], const double hi[3], double gradY,
                         double amp, const int patchKinds[6], const char* const patchNames[6])
{
    StructGen g{nb, nxb, ny, nz, Mapping{kind, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}, gradY, amp}, {}};
    g.cachePoints();
    return buildStructured(g, patchKinds, patchNames);
}

// sub-box [off, off+n) of a single-block global grid of gl[3] cells (same mapping => identical points as the global mesh)
void* icsmesh_structured_subbox(const int n[3], const int off[3], const int gl[3], int kind, const double lo[3], const double hi[3],
                                double gradY, double amp, const int patchKinds[6], const char* const patchNames[6])
{
    StructGen g{1, n[0], n[1], n[2], Mapping{kind, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}, gradY, amp}, {}};
    for (int d = 0; d < 3; d++) { g.off[d] = off[d]; g.glob[d] = gl[d]; }
    g.cachePoints();
    return buildStructured(g, patchKinds, patchNames);
}

void* icsmesh_read_polymesh(const char* dir) { return readPolyMesh(dir); }

void* icsmesh_extract_part(void* h, const int* part, int rank) { return extractPart(*(Mesh*)h, part, rank); }

void* icsmesh_renumber(void* h, const int* perm) { return renumber(*(Mesh*)h, perm); }

void icsmesh_free(void* h) { delete (Mesh*)h; }

const char* icsmesh_error(void* h) { return ((Mesh*)h)->error.c_str(); }

// sizes: nCells, nInternalFaces, nFaces, nPatches, solutionD[3]
void icsmesh_sizes(void* h, int out[7])
{
    Mesh& m = *(Mesh*)h;
    out[0] = m.nCells; out[1] = m.nInternalFaces; out[2] = m.nFaces; out[3] = (int)m.patches.size();
    out[4] = m.solutionD[0]; out[5] = m.solutionD[1]; out[6] = m.solutionD[2];
}

void icsmesh_patch(void* h, int i, int out[5], char name[64])
{
    Mesh& m = *(Mesh*)h;
    const Patch& p = m.patches[i];
    out[0] = p.kind; out[1] = p.start; out[2] = p.size; out[3] = p.nbrRank; out[4] = p.nbrPatch;
    std::snprintf(name, 64, "%s", p.name.c_str());
}

void icsmesh_set_patch_kind(void* h, int i, int kind, int nbrPatch)
{
    Mesh& m = *(Mesh*)h;
    m.patches[i].kind = kind;
    m.patches[i].nbrPatch = nbrPatch;
}

// coupled geometry of a translational cyclic pair (after both patches were given kind CYCLIC)
void icsmesh_cyclic_geometry(void* h, int ia, int ib) { cyclicGeometry(*(Mesh*)h, ia, ib); }

// copy arrays out (caller allocates): owner[nFaces], neighbour[nInternalFaces], Sf[3nFaces], Cf[3nFaces],
// magSf, weights, deltaCoeffs, nonOrthDeltaCoeffs [nFaces], C[3nCells], V[nCells]
void icsmesh_arrays(void* h, int* owner, int* neighbour, double* Sf, double* Cf, double* magSf, double* weights,
                    double* deltaCoeffs, double* nonOrthDeltaCoeffs, double* C, double* V)
{
    Mesh& m = *(Mesh*)h;
    std::memcpy(owner, m.owner.data(), sizeof(int) * m.nFaces);
    std::memcpy(neighbour, m.neighbour.data(), sizeof(int) * m.nInternalFaces);
    std::memcpy(Sf, m.Sf.data(), sizeof(V3) * m.nFaces);
    std::memcpy(Cf, m.Cf.data(), sizeof(V3) * m.nFaces);
    std::memcpy(magSf, m.magSf.data(), sizeof(double) * m.nFaces);
    std::memcpy(weights, m.weights.data(), sizeof(double) * m.nFaces);
    std::memcpy(deltaCoeffs, m.deltaCoeffs.data(), sizeof(double) * m.nFaces);
    std::memcpy(nonOrthDeltaCoeffs, m.nonOrthDeltaCoeffs.data(), sizeof(double) * m.nFaces);
    std::memcpy(C, m.C.data(), sizeof(V3) * m.nCells);
    std::memcpy(V, m.V.data(), sizeof(double) * m.nCells);
}

// partition bookkeeping (only for meshes made by icsmesh_extract_part)
void icsmesh_part_maps(void* h, int* cellGlobal, int* faceGlobal)
{
    Mesh& m = *(Mesh*)h;
    if (cellGlobal) std::memcpy(cellGlobal, m.cellGlobal.data(), sizeof(int) * m.cellGlobal.size());
    if (faceGlobal) std::memcpy(faceGlobal, m.faceGlobal.data(), sizeof(int) * m.faceGlobal.size());
}

}  // extern "C"
