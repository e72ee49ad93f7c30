import sys; sys.path.insert(0,'tests')
from _eng import hash_engine
e = hash_engine(4, board=9, sims=8)
print("created")
e.search(); print(e.root_stats(0)["N"][:5]); e.close()
