import importlib, sys
sys.path.insert(0, "/root/repo")
P = importlib.import_module("k8s-device-plugin_b200")
logs=[]
P._native.set_log_callback(lambda l,m: logs.append((l,m)))
with P.Context("cuda:devices=0,bytes=%d,prearm=1" % (96<<20)) as a:
    for step in range(5):
        r = a.probe_health(timed=False, min_gbs=1e-3)[0]
        print(step, hex(r.flags), r.seed, r.healthy, r.ms_device)
print(logs)
