import sys
sys.path.insert(0, ".")
src = open("tools/tc_conv_probe.py").read().split('if __name__ != "__main__"')[0]
exec(src)
names = {0: "full", 1: "no epilogue I/O", 2: "no MMA issue", 4: "no operand loads", 3: "no epi I/O + no MMA", 5: "no epi I/O + no loads", 6: "no MMA + no loads", 7: "barriers only"}
for nt, mt in ((128, 2), (64, 4)):
    print(f"C=256 L=12000 conv1 nt={nt} mt={mt}")
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        print(f"   {names[dbg]:24s} {probe(256, 12000, 0, dbg, nt=nt, mt=mt):8.1f} us")
print("C=128 L=12000 conv1 auto (same position count, narrower)")
for dbg in (0, 2, 4, 7):
    print(f"   {names[dbg]:24s} {probe(128, 12000, 0, dbg):8.1f} us")
