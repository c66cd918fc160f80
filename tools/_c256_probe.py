import sys
sys.path.insert(0, ".")
src = open("tools/tc_conv_probe.py").read().split('if __name__ != "__main__"')[0]
exec(src)
for Cc, L in ((256, 12000), (256, 10000), (128, 120000)):
    for conv2 in (0, 1):
        for nt, mt in ((0, 0), (128, 1), (128, 2), (64, 2), (64, 4), (32, 4)):
            try:
                t = probe(Cc, L, conv2, 0, nt=nt, mt=mt, blocked=3 if conv2 else 0)
            except Exception as e:
                t = float("nan")
            print(f"C={Cc} L={L} conv{conv2 + 1} nt={nt:3d} mt={mt}: {t:7.1f} us")
