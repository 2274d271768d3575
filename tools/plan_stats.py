import sys, os, json
sys.path.insert(0,'/root/repo')
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import plan_describe
n=int(sys.argv[1]); order=int(sys.argv[2]) if len(sys.argv)>2 else 1
K,w=oc.random_problem(n, 20260700+n)
plan=plan_describe(K,w,0.02,order,1)
tot_r=0; 
for i,p in enumerate(plan["passes"]):
    rs=p["rounds"]
    print(i, "shared",p["n_shared"],"fl",p["fuse_load"],"fs",p["fuse_store"],"phase",p["has_phase"],"gates",p["ngates"],"rounds",len(rs), [("S" if r.get("split") else "C")+str(len(r["gates"])) for r in rs])
    tot_r+=len(rs)
print("passes",len(plan["passes"]),"rounds",tot_r,"gates",plan["total_gates"])
