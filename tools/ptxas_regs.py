import re,subprocess,sys
for f in sys.argv[1:]:
    log=open(f).read()
    blocks=log.split("ptxas info    : Compiling entry function '")[1:]
    rows=[]
    for b in blocks:
        name=b.split("'")[0]
        regs=re.search(r'Used (\d+) registers',b).group(1)
        sp=re.search(r'(\d+) bytes spill stores',b).group(1)
        dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
        m=re.search(r'(scan_\w+_kernel)<(.*?)>',dem)
        rows.append((m.group(1),m.group(2),regs,sp))
    for r in sorted(rows): print(*r)
