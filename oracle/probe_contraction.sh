#!/bin/sh
# Evidence for oracle decision (i): how nvcc's default -fmad=true contracts the upstream
# pointnet2_ops distance expressions.  Prints the FMUL/FFMA order for sm_75 (the newest arch in
# upstream's TORCH_CUDA_ARCH_LIST) and sm_100a.  Observed with nvcc 12.9.86 for both:
#   a*a + b*b + c*c  ->  FMUL t = b*b ; FFMA t = a*a + t ; FFMA t = c*c + t
set -e
tmp=$(mktemp -d)
cat > "$tmp/contract.cu" <<'CU'
extern "C" __global__ void fps_expr(const float* p, float* out, float x1, float y1, float z1) {
  float x2 = p[0], y2 = p[1], z2 = p[2];
  float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
  float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
  out[0] = mag; out[1] = d; out[2] = (mag <= 1e-3) ? 1.f : 0.f;
}
extern "C" __global__ void bq_expr(const float* p, float* out, float new_x, float new_y, float new_z, float radius) {
  float radius2 = radius * radius;
  float x = p[0], y = p[1], z = p[2];
  float d2 = (new_x - x) * (new_x - x) + (new_y - y) * (new_y - y) + (new_z - z) * (new_z - z);
  out[0] = d2; out[1] = (d2 < radius2) ? 1.f : 0.f;
}
CU
for arch in "-arch=sm_75" "-gencode arch=compute_100a,code=sm_100a"; do
  echo "== nvcc $arch"
  nvcc $arch -O3 -cubin -o "$tmp/c.cubin" "$tmp/contract.cu"
  cuobjdump -sass "$tmp/c.cubin" | grep -E "Function|FMUL|FFMA|FADD|DSETP|FSETP|F2F"
done
rm -rf "$tmp"
