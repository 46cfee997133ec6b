import sys
sys.argv=[sys.argv[0]]
exec(open('tools/e2e_profile.py').read().split("print(prof.key_averages()")[0])
rows=[(e.key, e.count, e.self_device_time_total/3e3) for e in prof.key_averages() if e.self_device_time_total>0]
rows.sort(key=lambda r:-r[2])
tot=sum(r[2] for r in rows)
print("total device ms/step %.2f"%tot)
for k,c,t in rows[:45]: print("%-110s %4d %8.3f"%(k[:110],c//3,t))
