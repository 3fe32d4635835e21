import sys; sys.path.insert(0, "/root/repo")
import socialmediamodel_b200 as smm
ctx = smm.Context.default()
names = ["ex2", "sqrt", "rsqrt", "sqrt+ex2", "rsqrt+ex2", "lg2", "rcp", "sin"]
for op, nm in enumerate(names):
    print(nm, round(ctx.microbench("mufu", op), 1), "G MUFU/s")
