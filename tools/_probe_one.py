import sys
sys.path.insert(0, ".")
exec(open("tools/tc_conv_probe.py").read().split("names = {")[0])
Cc, L, conv2 = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
print(probe(Cc, L, conv2, 0, reps=3))
