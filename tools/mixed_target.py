import sys
sys.path.insert(0,".")
import drt_b200 as drt
with drt.Context(0) as ctx:
    ctx.upload(drt.cornell_box(1024,1024))
    for rep in range(2):
        img,grad,st=ctx.render(drt.make_opts(256,8,1.0,precision=drt.MIXED),stats=True)
    print(st.kernel_ms, st.retraced_paths/st.paths)
