// Small fp32 vector / quaternion / 3x3 helpers for the device code.
//
// Operation ORDER follows the GLM 0.9.9.8 formulas the reference is compiled against
// (reference vendor/glm 0.9.9.8/glm: detail/type_quat.inl:343-350 quat*vec3, :282-292 quat*quat,
// ext/quaternion_common.inl:119-122 inverse, ext/quaternion_geometric.inl:17-24 normalize,
// gtc/quaternion.inl:41-66 mat3_cast, detail/func_geometric.inl:48-90 dot/cross/normalize,
// detail/type_mat3x3.inl:468-520 mat*vec / mat*mat, detail/func_matrix.inl:269-291 inverse),
// so translation units built with -fmad=false reproduce the reference's fp32 results bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

#define PB_HD __host__ __device__ __forceinline__

struct V3 { float x, y, z; };
struct V2 { float x, y; };
struct Q4 { float x, y, z, w; };
struct M3 { V3 c[3]; };  // column-major like glm::mat3: c[i] is column i (basis vector i)

PB_HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PB_HD V3 mk3(float s) { return mk3(s, s, s); }
PB_HD V3 mk3(float4 v) { return mk3(v.x, v.y, v.z); }
PB_HD V2 mk2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
PB_HD Q4 mkq(float4 v) { Q4 q; q.x = v.x; q.y = v.y; q.z = v.z; q.w = v.w; return q; }
PB_HD float4 f4(V3 v, float w = 0.f) { return make_float4(v.x, v.y, v.z, w); }
PB_HD float4 f4(Q4 q) { return make_float4(q.x, q.y, q.z, q.w); }

PB_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
PB_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PB_HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PB_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PB_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
PB_HD V3 operator/(V3 a, V3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PB_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
PB_HD V3& operator-=(V3& a, V3 b) { a = a - b; return a; }
PB_HD V2 operator+(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
PB_HD V2 operator-(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
PB_HD V2 operator-(V2 a) { return mk2(-a.x, -a.y); }
PB_HD V2 operator*(float s, V2 a) { return mk2(s * a.x, s * a.y); }

PB_HD float get(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
PB_HD void set(V3& v, int i, float s) { if (i == 0) v.x = s; else if (i == 1) v.y = s; else v.z = s; }

PB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PB_HD V3 cross(V3 x, V3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
PB_HD float length2(V3 a) { return dot(a, a); }
PB_HD float length(V3 a) { return sqrtf(dot(a, a)); }
PB_HD float distance2(V3 a, V3 b) { return length2(b - a); }       // glm::distance2 = length2(p1 - p0)
PB_HD float distance(V3 a, V3 b) { return length(b - a); }
PB_HD V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
PB_HD V3 vmin(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PB_HD V3 vmax(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
PB_HD V3 vabs(V3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// glm::min(a,b) = (b < a) ? b : a ; glm::max(a,b) = (a < b) ? b : a  (differs from fminf only for NaN)
PB_HD float gmin(float a, float b) { return (b < a) ? b : a; }
PB_HD float gmax(float a, float b) { return (a < b) ? b : a; }
PB_HD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
PB_HD V3 gclamp(V3 v, V3 lo, V3 hi) { return mk3(gclamp(v.x, lo.x, hi.x), gclamp(v.y, lo.y, hi.y), gclamp(v.z, lo.z, hi.z)); }
PB_HD float gsign(float x) { return (float)((0.f < x) - (x < 0.f)); }
PB_HD float gmix(float a, float b, float t) { return a * (1.f - t) + b * t; }  // glm::mix for floats: x*(1-a) + y*a

// quaternion (x,y,z,w in memory, like glm::quat)
PB_HD V3 rotate(Q4 q, V3 v) {
    V3 qv = mk3(q.x, q.y, q.z);
    V3 uv = cross(qv, v);
    V3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
PB_HD Q4 qmul(Q4 p, Q4 q) {
    Q4 r;
    r.w = p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z;
    r.x = p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y;
    r.y = p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z;
    r.z = p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x;
    return r;
}
PB_HD float qdot(Q4 a, Q4 b) { return (a.w * b.w + a.x * b.x) + (a.y * b.y + a.z * b.z); }
PB_HD Q4 qinverse(Q4 q) {
    float d = qdot(q, q);
    Q4 r; r.w = q.w / d; r.x = -q.x / d; r.y = -q.y / d; r.z = -q.z / d;
    return r;
}
PB_HD Q4 qnormalize(Q4 q) {
    float len = sqrtf(qdot(q, q));
    Q4 r;
    if (len <= 0.f) { r.x = 0.f; r.y = 0.f; r.z = 0.f; r.w = 1.f; return r; }
    float inv = 1.0f / len;
    r.w = q.w * inv; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv;
    return r;
}
PB_HD M3 mat3_cast(Q4 q) {
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    M3 m;
    m.c[0] = mk3(1.f - 2.f * (qyy + qzz), 2.f * (qxy + qwz), 2.f * (qxz - qwy));
    m.c[1] = mk3(2.f * (qxy - qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz + qwx));
    m.c[2] = mk3(2.f * (qxz + qwy), 2.f * (qyz - qwx), 1.f - 2.f * (qxx + qyy));
    return m;
}

// 3x3
PB_HD V3 mul(const M3& m, V3 v) {
    return mk3(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z,
               m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
               m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z);
}
// v * m  (== transpose(m) * v), glm row-vector product
PB_HD V3 mulT(const M3& m, V3 v) { return mk3(dot(m.c[0], v), dot(m.c[1], v), dot(m.c[2], v)); }
PB_HD M3 transpose(const M3& m) {
    M3 r;
    r.c[0] = mk3(m.c[0].x, m.c[1].x, m.c[2].x);
    r.c[1] = mk3(m.c[0].y, m.c[1].y, m.c[2].y);
    r.c[2] = mk3(m.c[0].z, m.c[1].z, m.c[2].z);
    return r;
}
PB_HD M3 mul(const M3& a, const M3& b) {
    M3 r;
    r.c[0] = mul(a, b.c[0]);
    r.c[1] = mul(a, b.c[1]);
    r.c[2] = mul(a, b.c[2]);
    return r;
}
PB_HD M3 operator+(const M3& a, const M3& b) { M3 r; r.c[0] = a.c[0] + b.c[0]; r.c[1] = a.c[1] + b.c[1]; r.c[2] = a.c[2] + b.c[2]; return r; }
PB_HD M3 operator-(const M3& a, const M3& b) { M3 r; r.c[0] = a.c[0] - b.c[0]; r.c[1] = a.c[1] - b.c[1]; r.c[2] = a.c[2] - b.c[2]; return r; }
PB_HD M3 operator*(float s, const M3& a) { M3 r; r.c[0] = a.c[0] * s; r.c[1] = a.c[1] * s; r.c[2] = a.c[2] * s; return r; }
PB_HD float m(const M3& a, int col, int row) { return get(a.c[col], row); }
PB_HD float det3(const M3& a) {
    return + a.c[0].x * (a.c[1].y * a.c[2].z - a.c[2].y * a.c[1].z)
           - a.c[1].x * (a.c[0].y * a.c[2].z - a.c[2].y * a.c[0].z)
           + a.c[2].x * (a.c[0].y * a.c[1].z - a.c[1].y * a.c[0].z);
}
PB_HD M3 inverse(const M3& a) {
    float inv = 1.0f / det3(a);
    M3 r;
    r.c[0].x = +(a.c[1].y * a.c[2].z - a.c[2].y * a.c[1].z) * inv;
    r.c[1].x = -(a.c[1].x * a.c[2].z - a.c[2].x * a.c[1].z) * inv;
    r.c[2].x = +(a.c[1].x * a.c[2].y - a.c[2].x * a.c[1].y) * inv;
    r.c[0].y = -(a.c[0].y * a.c[2].z - a.c[2].y * a.c[0].z) * inv;
    r.c[1].y = +(a.c[0].x * a.c[2].z - a.c[2].x * a.c[0].z) * inv;
    r.c[2].y = -(a.c[0].x * a.c[2].y - a.c[2].x * a.c[0].y) * inv;
    r.c[0].z = +(a.c[0].y * a.c[1].z - a.c[1].y * a.c[0].z) * inv;
    r.c[1].z = -(a.c[0].x * a.c[1].z - a.c[1].x * a.c[0].z) * inv;
    r.c[2].z = +(a.c[0].x * a.c[1].y - a.c[1].x * a.c[0].y) * inv;
    return r;
}
// glm::matrixCross3(x): skew-symmetric matrix, Result[0][1]=x.z, [0][2]=-x.y, [1][0]=-x.z, [1][2]=x.x, [2][0]=x.y, [2][1]=-x.x
PB_HD M3 matrixCross3(V3 x) {
    M3 r;
    r.c[0] = mk3(0.f, x.z, -x.y);
    r.c[1] = mk3(-x.z, 0.f, x.x);
    r.c[2] = mk3(x.y, -x.x, 0.f);
    return r;
}
PB_HD float det2(V2 c0, V2 c1) { return c0.x * c1.y - c1.x * c0.y; }  // glm::determinant(mat2(c0,c1))
