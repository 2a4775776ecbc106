import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and 'rn_assign' in r[4]]
v=[float(r[-1])/1e3 for r in rows]
names=["coco_mixed","coco_empty","coco_full","pascal_mixed","pascal_empty","pascal_full"]
print({n: round(min(v[3*i:3*i+3]),1) for i,n in enumerate(names)})
