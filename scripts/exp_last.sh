run() {
  B200DR_LIB=$2 timeout -k 5 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $3 > gpurun_out/b_var.json 2>gpurun_out/b_var.err
  python -c "
import json; d=json.load(open('gpurun_out/b_var.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if 'grad' in k))" || tail -3 gpurun_out/b_var.err
}
run base "" ""
run gacc $PWD/nvdiffrast_b200/variants/libb200dr_gacc.so ""
