"""development: prints the edge backward kernel's timeline stamps (stderr of a run with HGN_TC_ABLATE bit 64) relative to the previous tile's E5"""
import sys
lines=[l for l in open(sys.argv[1]) if l.startswith('tile')][-8:]
T=[[int(x) for x in l.split(':')[1].split()] for l in lines]
keys=[int(k) for k in sys.argv[2].split(',')] if len(sys.argv)>2 else [0,1,2,10,11,3,12,13,4,14,15,5,16,17,6,18,19,7,20,8,21]
print('      '+' '.join(f'{k:6d}' for k in keys))
for t in range(1,8):
    base=T[t-1][20]
    print(f't{t} '+' '.join(f'{T[t][k]-base:6d}' if T[t][k]>=0 else '     -' for k in keys), ' period', T[t][20]-T[t-1][20])
