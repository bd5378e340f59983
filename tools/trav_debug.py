import sys
sys.path.insert(0,'.')
import drt_b200 as drt
scene = drt.tessellated_room(204, 362, width=512, height=512)
with drt.Context(0) as ctx:
    ctx.upload(scene)
    for B in (1, 2, 8):
        img, grad, st = ctx.render(drt.make_opts(16, B, 1.0), stats=True)
        seg = st.segments
        print(f"B={B} segs {seg} nodes/seg {st.bvh_nodes/seg:.2f} tests/seg {st.tri_tests/seg:.2f} stale/seg {st.truncated_paths/seg:.2f} "
              f"inner-hits/node {st.retraced_paths/st.bvh_nodes:.2f} leaf-tri-hits/node {st.paths/st.bvh_nodes:.2f} ms {st.kernel_ms:.1f}")
