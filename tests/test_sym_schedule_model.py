"""Model of csrc/sym_kernel.cu's control logic — tile geometry (packed / strided / pair tiles), the stage
schedule of the radix-4 register stages (single-level, two-level and centre stages), the pass plan of
`extend_sym` and the fused-combine epilogue — in plain Python over a small prime field, compared with the
straightforward level-by-level butterfly network and combine (reference src/fftree.rs:72-120, 155-159 in
symmetric form).  It mirrors the CUDA code statement by statement so that an index or schedule change can
be tried on the CPU first; the GPU parity tests remain the check of the CUDA code itself."""
import random
P=1000003
rnd=random.Random(3)
def ins0(q,b): return ((q>>b)<<(b+1))|(q&((1<<b)-1))
def kernel(p, mem):
    T=1<<p['log_t']; hmask=(1<<p['log_h'])-1
    tiles=(p['total']+T-1)>>p['log_t']
    for blk in range(tiles):
        if p['packed']:
            tm=dict(log_c=p['log_t'],cmask=T-1,rmask=0,krows=0,row_shift=0,log_h=p['log_h'],pos0=0); gbase=blk<<p['log_t']
        else:
            w=blk%p['nv']; tile=blk//p['nv']; ncg=p['row_shift']-p['log_c']
            cg=tile&((1<<ncg)-1); q_hi=tile>>ncg
            pos0=(q_hi<<p['lvl_hi'])+(cg<<p['log_c'])
            tm=dict(log_c=p['log_c'],cmask=(1<<p['log_c'])-1,rmask=(1<<p['krows'])-1,krows=p['krows'],row_shift=p['row_shift'],log_h=p['log_h'],pos0=pos0)
            gbase=((2*w if p['pair'] else w)<<p['log_h'])+pos0
        def mp(e):
            rr=e>>tm['log_c']; c=e&tm['cmask']; rel=((rr&tm['rmask'])<<tm['row_shift'])+c
            return ((rr>>tm['krows'])<<tm['log_h'])+rel, tm['pos0']+rel
        s=[0]*T
        for e in range(T):
            go,pos=mp(e); g=gbase+go
            s[e]=mem[p['in']][g] if g<p['total'] else 0
        nlev=p['lvl_hi']-p['lvl_lo']; mid=p['do_d'] and p['do_r'] and nlev>=2
        j_base=p['lvl_lo']+(2 if mid else 0); cnt=p['lvl_hi']-j_base; odd=cnt&1; npairs=cnt>>1
        nD=odd+npairs if p['do_d'] else 0; nR=odd+npairs if p['do_r'] else 0
        nst=nD+(1 if mid else 0)+nR
        outbuf={}
        for sidx in range(nst):
            if sidx<nD:
                if odd and sidx==0: ops,jh,jl=1,p['lvl_hi']-1,0
                else:
                    u=sidx-odd; ops=3; jh=p['lvl_hi']-1-odd-2*u; jl=jh-1
            elif mid and sidx==nD: ops,jh,jl=1|16|8,p['lvl_lo']+1,p['lvl_lo']
            else:
                t=sidx-nD-(1 if mid else 0)
                if t<npairs: ops=12; jl=j_base+2*t; jh=jl+1
                else: ops,jh,jl=8,p['lvl_hi']-1,0
            two=(ops&(2|4|16))!=0
            b_hi=(jh+p['boff'])&0xffffffff
            b_lo=((jl+p['boff'])&0xffffffff) if two else (1 if b_hi==0 else b_hi-1)
            b1,b2=min(b_lo,b_hi),max(b_lo,b_hi)
            mh,ml=(1<<jh)-1,(1<<jl)-1
            first=sidx==0 and p['pre'] is not None
            last=sidx+1==nst; to_global=last and not p['comb']
            D,R=mem[p['tw_d']],mem[p['tw_r']]
            for q in range(T//4):
                e0=ins0(ins0(q,b1),b2); e1=e0+(1<<b_lo); e2=e0+(1<<b_hi); e3=e1+(1<<b_hi)
                (g0,pa),(g1,pb),(g2,pc),(g3,pd)=mp(e0),mp(e1),mp(e2),mp(e3)
                x=[s[e0],s[e1],s[e2],s[e3]]
                if first:
                    pre=mem[p['pre']]
                    x=[x[0]*pre[pa&hmask]%P,x[1]*pre[pb&hmask]%P,x[2]*pre[pc&hmask]%P,x[3]*pre[pd&hmask]%P]
                def dp(a,b,g): return (a+b)%P,(a-b)*g%P
                def rp(a,b,g):
                    t=g*b%P; return (a+t)%P,(a-t)%P
                if ops&1:
                    x[0],x[2]=dp(x[0],x[2],D[(1<<jh)+(pa&mh)]); x[1],x[3]=dp(x[1],x[3],D[(1<<jh)+(pb&mh)])
                if ops&2:
                    g=D[(1<<jl)+(pa&ml)]; x[0],x[1]=dp(x[0],x[1],g); x[2],x[3]=dp(x[2],x[3],g)
                if ops&16:
                    c=mem['ctr']
                    def cp(a,b):
                        t=c*(a-b)%P; s_=(a+b)%P; return (s_+t)%P,(s_-t)%P
                    x[0],x[1]=cp(x[0],x[1]); x[2],x[3]=cp(x[2],x[3])
                if ops&4:
                    g=R[(1<<jl)+(pa&ml)]; x[0],x[1]=rp(x[0],x[1],g); x[2],x[3]=rp(x[2],x[3],g)
                if ops&8:
                    x[0],x[2]=rp(x[0],x[2],R[(1<<jh)+(pa&mh)]); x[1],x[3]=rp(x[1],x[3],R[(1<<jh)+(pb&mh)])
                if to_global:
                    if p['post'] is not None:
                        po=mem[p['post']]
                        x=[x[0]*po[pa&hmask]%P,x[1]*po[pb&hmask]%P,x[2]*po[pc&hmask]%P,x[3]*po[pd&hmask]%P]
                    for gg,xx in zip((g0,g1,g2,g3),x):
                        if gbase+gg<p['total']: outbuf[gbase+gg]=xx
                else:
                    s[e0],s[e1],s[e2],s[e3]=x
        for g,v in outbuf.items(): mem[p['out']][g]=v
        if p['comb']:
            ush=p['log_h'] if p['packed'] else p['log_t']-1
            for idx in range(T//2):
                eu=((idx>>ush)<<(ush+1))|(idx&((1<<ush)-1)); ev=eu+(1<<ush)
                (gu,pu),(gv,pv)=mp(eu),mp(ev); gu+=gbase; gv+=gbase
                if gv>=p['total']: continue
                i=pu&hmask; o=(gu-i)+2*i
                A=mem[p['A']]
                mem[p['out']][o]=(A[gu]+A[gv]*mem['xnn'][2*i])%P
                mem[p['out']][o+1]=(mem['gam'][i]*s[eu]+mem['gx'][i]*s[ev])%P
def extend_sym(mem,inn,out,log_h,nvec,pre,post,comb,LT):
    total=nvec<<log_h
    if total<4 or log_h==0: return False
    if comb and nvec&1: return False
    base=dict(tw_d='twd',tw_r='twr',total=total,log_h=log_h,A=(comb['A'] if comb else None),pair=0,comb=0,nv=1,log_c=0,krows=0,row_shift=0,packed=0,pre=None,post=None)
    need=log_h+(1 if comb else 0)
    if need<=LT:
        p=dict(base); p.update({'in':inn,'out':comb['out'] if comb else out,'packed':1,'log_t':max(need,2)})
        while p['log_t']<LT and (1<<p['log_t'])<total: p['log_t']+=1
        p.update(lvl_lo=0,lvl_hi=log_h,boff=0,do_d=1,do_r=1,pre=pre,post=None if comb else post,comb=1 if comb else 0)
        kernel(p,mem); return True
    if log_h<LT: return False
    outer=log_h-LT
    if comb and outer==0: return False
    kmax=LT-5 if LT>5 else 1
    npass=(outer+kmax-1)//kmax
    bounds=[log_h]+[log_h-(outer*i)//npass for i in range(1,npass+1)]
    src=inn
    for i in range(npass):
        p=dict(base); p.update({'in':src,'out':out,'log_t':LT,'lvl_hi':bounds[i],'lvl_lo':bounds[i+1]})
        p['krows']=p['lvl_hi']-p['lvl_lo']; p['log_c']=LT-p['krows']; p['row_shift']=p['lvl_lo']; p['boff']=(p['log_c']-p['lvl_lo'])&0xffffffff
        p.update(nv=nvec,do_d=1,do_r=0,pre=pre if i==0 else None)
        kernel(p,mem); src=out
    p=dict(base); p.update({'in':src,'out':out,'log_t':LT,'packed':1,'lvl_lo':0,'lvl_hi':LT,'boff':0,'do_d':1,'do_r':1,'pre':pre if npass==0 else None,'post':post if npass==0 else None})
    kernel(p,mem)
    for i in range(npass-1,-1,-1):
        fin=i==0
        p=dict(base); pair=1 if (fin and comb) else 0
        p.update({'in':out,'out':comb['out'] if pair else out,'log_t':LT,'pair':pair,'comb':pair,'lvl_hi':bounds[i],'lvl_lo':bounds[i+1]})
        p['krows']=p['lvl_hi']-p['lvl_lo']; p['log_c']=LT-pair-p['krows']; p['row_shift']=p['lvl_lo']; p['boff']=(p['log_c']-p['lvl_lo'])&0xffffffff
        p.update(nv=nvec//2 if pair else nvec,do_d=0,do_r=1,post=post if (fin and not comb) else None)
        kernel(p,mem)
    return True
def ref_extend(x,log_h,D,R,pre,post):
    h=1<<log_h; x=[a*pre[i]%P for i,a in enumerate(x)] if pre else list(x)
    for j in range(log_h-1,-1,-1):
        for pp in range(h):
            if not (pp>>j)&1:
                q=pp+(1<<j); g=D[(1<<j)+(pp&((1<<j)-1))]; a,b=x[pp],x[q]
                x[pp]=(a+b)%P; x[q]=(a-b)*g%P
    for j in range(log_h):
        for pp in range(h):
            if not (pp>>j)&1:
                q=pp+(1<<j); g=R[(1<<j)+(pp&((1<<j)-1))]; a,b=x[pp],x[q]; t=g*b%P
                x[pp]=(a+t)%P; x[q]=(a-t)%P
    return [a*post[i]%P for i,a in enumerate(x)] if post else x
def test_schedule_model_matches_the_level_by_level_network():
  ok=0
  for LT in (6,7):
    for log_h in range(1,LT+7):
      h=1<<log_h
      for nvec in (1,2,4):
        if (nvec<<log_h)>(1<<15): continue
        mem={'twd':[rnd.randrange(1,P) for _ in range(h)],'twr':[rnd.randrange(1,P) for _ in range(h)],
             'pre':[rnd.randrange(1,P) for _ in range(h)],'post':[rnd.randrange(1,P) for _ in range(h)],
             'xnn':[rnd.randrange(1,P) for _ in range(2*h)],'gam':[rnd.randrange(1,P) for _ in range(h)],'gx':[rnd.randrange(1,P) for _ in range(h)]}
        mem['ctr']=mem['twr'][1]*mem['twd'][1]%P if h>=2 else 0
        x=[rnd.randrange(P) for _ in range(nvec*h)]
        for usepre,usepost in ((1,1),(0,0),(1,0)):
          mem['in']=list(x); mem['out']=[None]*(nvec*h)
          r=extend_sym(mem,'in','out',log_h,nvec,'pre' if usepre else None,'post' if usepost else None,None,LT)
          want=[]
          for v in range(nvec): want+=ref_extend(x[v*h:(v+1)*h],log_h,mem['twd'],mem['twr'],mem['pre'] if usepre else None,mem['post'] if usepost else None)
          if r: assert mem['out']==want,(LT,log_h,nvec,usepre,usepost); ok+=1
          else: assert (nvec<<log_h)<4
        # in-place extend
        mem['in']=list(x); 
        r=extend_sym(mem,'in','in',log_h,nvec,'pre',None,None,LT)
        if r:
          want=[]
          for v in range(nvec): want+=ref_extend(x[v*h:(v+1)*h],log_h,mem['twd'],mem['twr'],mem['pre'],None)
          assert mem['in']==want,("inplace",LT,log_h,nvec); ok+=1
        if nvec>=2:
          mem['in']=list(x); mem['W']=[None]*(nvec*h); mem['dst']=[None]*(nvec*h)
          r=extend_sym(mem,'in','W',log_h,nvec,'pre',None,dict(A='in',out='dst'),LT)
          W=[]
          for v in range(nvec): W+=ref_extend(x[v*h:(v+1)*h],log_h,mem['twd'],mem['twr'],mem['pre'],None)
          want=[None]*(nvec*h)
          for b in range(nvec//2):
              off=b*2*h
              for i in range(h):
                  want[off+2*i]=(x[off+i]+x[off+h+i]*mem['xnn'][2*i])%P
                  want[off+2*i+1]=(mem['gam'][i]*W[off+i]+mem['gx'][i]*W[off+h+i])%P
          if r: assert mem['dst']==want,("comb",LT,log_h,nvec); ok+=1
          else: assert log_h==LT,("comb refused",LT,log_h,nvec)
  assert ok > 300


def test_tma_box_schedule_covers_every_tile_element_once():
    """The TMA tile load of sym_tile<.., TMA = true> (one thread issues cp.async.bulk.tensor.3d boxes of {4 limbs,
    2^min(log_c, 8) columns, 2^krows rows — one row when a row is wider than a box}; the tensor is
    [element >> rs][element & (2^rs - 1)][8 limbs]) must put the low / high half of every tile element exactly where the
    cp.async loop puts it: shared-memory slot e <- global element gbase + goff(e) of TileMap.  Index arithmetic only, for
    every pass geometry the planner produces at the real tile size (packed, strided and pair tiles, log_h up to 22)."""
    global kernel
    passes = []
    saved = kernel
    kernel = lambda p, mem: passes.append(dict(p))
    try:
        LT = 10
        for log_h in range(1, 23):
            for nvec, comb in ((1, None), (2, None), (2, dict(A='in', out='dst')), (8, dict(A='in', out='dst'))):
                if (nvec << log_h) < 1024:
                    continue
                extend_sym({}, 'in', 'out', log_h, nvec, 'pre', 'post', comb, LT)
    finally:
        kernel = saved
    assert len(passes) > 150
    seen_shapes = set()
    for p in passes:
        T = 1 << p['log_t']
        if p['log_t'] < 9 or p['total'] % T:
            continue                                       # make_tile_map refuses: cp.async path
        rs = p['log_t'] if p['packed'] else p['row_shift']
        lc = p['log_t'] if p['packed'] else p['log_c']
        kr = 0 if p['packed'] else p['krows']
        assert p['total'] % (1 << rs) == 0
        bc_log = min(lc, 8)
        krb = 0 if lc > 8 else kr
        seen_shapes.add((p['packed'], p['pair'], lc, kr))
        tiles = p['total'] >> p['log_t']
        for blk in sorted({0, 1, tiles // 2, tiles - 1}):
            if blk >= tiles:
                continue
            if p['packed']:
                tm = dict(log_c=p['log_t'], cmask=T - 1, rmask=0, krows=0, row_shift=0, log_h=p['log_h'])
                gbase = blk << p['log_t']
            else:
                w, tile = blk % p['nv'], blk // p['nv']
                ncg = p['row_shift'] - p['log_c']
                pos0 = ((tile >> ncg) << p['lvl_hi']) + ((tile & ((1 << ncg) - 1)) << p['log_c'])
                tm = dict(log_c=p['log_c'], cmask=(1 << p['log_c']) - 1, rmask=(1 << p['krows']) - 1, krows=p['krows'],
                          row_shift=p['row_shift'], log_h=p['log_h'])
                gbase = ((2 * w if p['pair'] else w) << p['log_h']) + pos0
            want = []
            for e in range(T):
                rr, c = e >> tm['log_c'], e & tm['cmask']
                want.append(gbase + ((rr >> tm['krows']) << tm['log_h']) + ((rr & tm['rmask']) << tm['row_shift']) + c)
            got = [[None] * T, [None] * T]
            c1, c2 = gbase & ((1 << rs) - 1), gbase >> rs
            nchunk, nrow, npair = 1 << (lc - bc_log), 1 << (kr - krb), 2 if p['pair'] else 1
            nbytes = 0
            for hf in range(2):
                for pb in range(npair):
                    for rw in range(nrow):
                        for cc in range(nchunk):
                            dst = (pb << (kr + lc)) + (rw << lc) + (cc << bc_log)          # element offset inside this half
                            prow = (1 << (p['log_h'] - rs)) if p['pair'] else 0
                            x1, x2 = c1 + (cc << bc_log), c2 + pb * prow + rw
                            for r in range(1 << krb):                                      # dense box: rows of 2^bc_log columns
                                for c in range(1 << bc_log):
                                    slot = dst + (r << bc_log) + c
                                    assert got[hf][slot] is None, ("written twice", p, blk, slot)
                                    got[hf][slot] = ((x2 + r) << rs) + x1 + c
                                    nbytes += 16
            assert got[0] == want and got[1] == want, (p, blk)
            assert nbytes == T * 32                                                        # the mbarrier's expected byte count
    assert any(lc > 8 and kr > 0 for _, _, lc, kr in seen_shapes)      # the shape whose rows are wider than a box
    assert any(pair for _, pair, _, _ in seen_shapes)
