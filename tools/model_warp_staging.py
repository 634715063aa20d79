#!/usr/bin/env python3
"""How often does a warp's gather footprint fit its shared-memory slice? A numpy model of k_render_warp's staging
decision on the BASELINE configs (window origins from the projection formulas, no rasters needed), for 32x1 strips
against 16x2 / 8x4 patches per warp and the 32x8 block tile, and for several budgets. It is what the 8x4 patch
and the 768-float slice in envutil_b200/csrc/render_impl.cuh were chosen from (CPU only, approximate at seams)."""
import numpy as np
def cube_pick(ray, face_px, section, frame, biatan=False):
    x,y,z = ray
    ax,ay,az = np.abs(x),np.abs(y),np.abs(z)
    m1 = ax>=ay; m2 = ax>=az; m3 = ay>=az
    xd = m1&m2; zd = (~m2)&(~m3); yd = ~(xd|zd)
    face = np.zeros(x.shape,int); u=np.zeros(x.shape); v=np.zeros(x.shape)
    # x dominant: RIGHT(1)/LEFT(0): in=(-z/x, y/|x|)
    f = np.where(x>0,1,0); face=np.where(xd,f,face); u=np.where(xd,-z/np.where(xd,x,1),u); v=np.where(xd,y/np.where(xd,ax,1),v)
    f = np.where(z>0,4,5); face=np.where(zd,f,face); u=np.where(zd,x/np.where(zd,z,1),u); v=np.where(zd,y/np.where(zd,az,1),v)
    f = np.where(y>0,3,2); face=np.where(yd,f,face); u=np.where(yd,-x/np.where(yd,ay,1),u); v=np.where(yd,z/np.where(yd,y,1),v)
    if biatan: u=4/np.pi*np.arctan(u); v=4/np.pi*np.arctan(v)
    px = frame + (u+1)/2*face_px - 0.5
    py = face*section + frame + (v+1)/2*face_px - 0.5
    return px, py
def sph_rays(W,H,rows):
    lon = (np.arange(W)+0.5)/W*2*np.pi-np.pi
    out=[]
    for yy in rows:
        lat = (yy+0.5)/H*np.pi-np.pi/2
        out.append((np.cos(lat)*np.sin(lon), np.full(W,np.sin(lat)), np.cos(lat)*np.cos(lon)))
    return out
def stats(name, coords_per_row, order, TS=3, maxf=1024, maxrows=16):
    tot=st=0; sizes=[]
    for px,py in coords_per_row:
        ix=np.floor(px).astype(int)-(order-1)//2; iy=np.floor(py).astype(int)-(order-1)//2
        n=len(ix)//32*32
        ix=ix[:n].reshape(-1,32); iy=iy[:n].reshape(-1,32)
        mnx=ix.min(1); mxx=ix.max(1); mny=iy.min(1); mxy=iy.max(1)
        a0=(mnx*TS)&~3; wf=(((mxx+order)*TS-a0)+3)&~3; rows=mxy-mny+order
        ok=(rows<=maxrows)&(rows*wf<=maxf)
        tot+=len(ok); st+=ok.sum(); sizes+=list((rows*wf)[ok])
    print("%-28s warps %7d staged %.3f  median floats %d p90 %d" % (name,tot,st/tot,np.median(sizes) if sizes else 0,np.percentile(sizes,90) if sizes else 0))
rng=np.random.default_rng(1)
# C2: cubemap 2048 (section 2112, frame 32) -> sph 8192x4096 cubic
rows=np.sort(rng.integers(0,4096,200))
stats("C2 cubic", [cube_pick(r,2048,2112,32) for r in sph_rays(8192,4096,rows)], 4)
for mf in (768,1536,2048): stats("C2 cubic maxf=%d"%mf, [cube_pick(r,2048,2112,32) for r in sph_rays(8192,4096,rows)], 4, maxf=mf)
# C3b: biatan6 4096 (section 4160, frame 32) -> sph 16384x8192 bilinear
rows=np.sort(rng.integers(0,8192,150))
stats("C3b bilinear", [cube_pick(r,4096,4160,32,True) for r in sph_rays(16384,8192,rows)], 2)
# C3a: latlon 16384x8192 -> biatan6 4096 faces, bilinear: target px (x, y in 6*4096)
def ba6_rows(F, ys):
    out=[]
    xs=(np.arange(F)+0.5)/F*2-1
    for y in ys:
        face=y//F; v=((y%F)+0.5)/F*2-1
        u=np.tan(xs*np.pi/4); vv=np.tan(v*np.pi/4)*np.ones(F)
        one=np.ones(F)
        ray={0:(-one,vv,u),1:(one,vv,-u),2:(-u,-one,-vv),3:(-u,one,vv),4:(u,vv,one),5:(-u,vv,-one)}[face]
        x,yv,z=ray
        lon=np.arctan2(x,z); lat=np.arctan2(yv,np.sqrt(x*x+z*z))
        out.append(((lon+np.pi)/(2*np.pi)*16384-0.5+2,(lat+np.pi/2)/np.pi*8192-0.5+2))
    return out
ys=np.sort(rng.integers(0,6*4096,200))
stats("C3a bilinear", ba6_rows(4096,ys), 2)
# C1: latlon 4096 -> rect 1920x1080 hfov 90
def rect_rows(W,H,hfov,ys,SW,SH):
    t=np.tan(np.radians(hfov)/2); xs=((np.arange(W)+0.5)/W*2-1)*t
    out=[]
    for y in ys:
        yv=((y+0.5)/H*2-1)*t*H/W*np.ones(W); z=np.ones(W)
        lon=np.arctan2(xs,z); lat=np.arctan2(yv,np.sqrt(xs*xs+z*z))
        out.append(((lon+np.pi)/(2*np.pi)*SW-0.5+2,(lat+np.pi/2)/np.pi*SH-0.5+2))
    return out
stats("C1 bilinear", rect_rows(1920,1080,90,np.arange(0,1080,7),4096,2048), 2)
# C4: latlon 8192x4096 -> fisheye 4096^2 hfov 180 (centre rays; twine margin mx=(bw*5)//64+2,my=2 added)
def fish_rows(W,ys,SW,SH):
    xs=((np.arange(W)+0.5)/W*2-1)*np.pi/2
    out=[]
    for y in ys:
        yv=((y+0.5)/W*2-1)*np.pi/2
        r=np.sqrt(xs*xs+yv*yv); a=np.pi/2-r; b=np.arctan2(xs,yv)
        z=np.sin(a); rr=np.cos(a); X=rr*np.sin(b); Y=rr*np.cos(b)
        lon=np.arctan2(X,z); lat=np.arctan2(Y,np.sqrt(X*X+z*z))
        out.append(((lon+np.pi)/(2*np.pi)*SW-0.5+2,(lat+np.pi/2)/np.pi*SH-0.5+2))
    return out
def stats_tw(name, coords, order, TS=3, maxf=1024, maxrows=16):
    tot=st=0
    for px,py in coords:
        ix=np.floor(px).astype(int); iy=np.floor(py).astype(int)
        n=len(ix)//32*32; ix=ix[:n].reshape(-1,32); iy=iy[:n].reshape(-1,32)
        mnx=ix.min(1); mxx=ix.max(1); mny=iy.min(1); mxy=iy.max(1)
        bw=mxx-mnx+1; mx=(bw*5)//64+2
        bx0=mnx-mx; bx1=mxx+mx; by0=mny-2; by1=mxy+2
        a0=(bx0*TS)&~3; wf=(((bx1+order)*TS-a0)+3)&~3; rows=by1-by0+order
        ok=(rows<=maxrows)&(rows*wf<=maxf); tot+=len(ok); st+=ok.sum()
    print("%-28s warps %7d staged %.3f" % (name,tot,st/tot))
ys=np.sort(rng.integers(0,4096,200))
stats_tw("C4 bilinear twine", fish_rows(4096,ys,8192,4096), 2)
stats_tw("C4 bilinear twine maxf=2048", fish_rows(4096,ys,8192,4096), 2, maxf=2048)

print("---- patch shapes (C2 cubic / C3a / C3b bilinear / C4 twine) ----")
def stats_patch(name, coordfn, H, order, pw, ph, TS=3, maxf=1024, maxrows=16, twine=False, nblk=150):
    tot=st=0; fl=[]
    ys0=np.sort(rng.integers(0,H//ph,nblk))*ph
    for y0 in ys0:
        rows=coordfn(np.arange(y0,y0+ph))
        PX=np.stack([r[0] for r in rows]); PY=np.stack([r[1] for r in rows])   # ph x W
        W=PX.shape[1]//pw*pw
        ix=np.floor(PX[:,:W]).astype(int)-(order-1)//2; iy=np.floor(PY[:,:W]).astype(int)-(order-1)//2
        ix=ix.reshape(ph,-1,pw).transpose(1,0,2).reshape(-1,ph*pw); iy=iy.reshape(ph,-1,pw).transpose(1,0,2).reshape(-1,ph*pw)
        mnx=ix.min(1); mxx=ix.max(1); mny=iy.min(1); mxy=iy.max(1)
        if twine:
            bw=mxx-mnx+1; mx=(bw*5)//64+2; mnx=mnx-mx; mxx=mxx+mx; mny=mny-2; mxy=mxy+2
        a0=(mnx*TS)&~3; wf=(((mxx+order)*TS-a0)+3)&~3; nr=mxy-mny+order
        ok=(nr<=maxrows)&(nr*wf<=maxf); tot+=len(ok); st+=ok.sum(); fl+=list((nr*wf)[ok])
    print("%-22s %2dx%d staged %.3f  mean floats %4d (%.1f floats/px)" % (name,pw,ph,st/tot,np.mean(fl),np.mean(fl)/(pw*ph)))
c2=lambda ys:[cube_pick(r,2048,2112,32) for r in sph_rays(8192,4096,ys)]
c3b=lambda ys:[cube_pick(r,4096,4160,32,True) for r in sph_rays(16384,8192,ys)]
c3a=lambda ys:ba6_rows(4096,ys)
c4=lambda ys:fish_rows(4096,ys,8192,4096)
for pw,ph in ((32,1),(16,2),(8,4),(32,8)):
    mf = 6144 if (pw,ph)==(32,8) else 1024
    mr = 256 if (pw,ph)==(32,8) else 16
    stats_patch("C2 cubic",c2,4096,4,pw,ph,maxf=mf,maxrows=mr)
    stats_patch("C3b bilinear",c3b,8192,2,pw,ph,maxf=mf,maxrows=mr)
    stats_patch("C3a bilinear",c3a,6*4096,2,pw,ph,maxf=mf,maxrows=mr)
    stats_patch("C4 bilinear twine",c4,4096,2,pw,ph,maxf=mf,maxrows=mr,twine=True)
print("---- 8x4 with smaller budgets ----")
for mf in (512, 640, 768):
    stats_patch("C2 cubic mf=%d"%mf,c2,4096,4,8,4,maxf=mf)
    stats_patch("C3b bilinear mf=%d"%mf,c3b,8192,2,8,4,maxf=mf)
    stats_patch("C3a bilinear mf=%d"%mf,c3a,6*4096,2,8,4,maxf=mf)
    stats_patch("C4 twine mf=%d"%mf,c4,4096,2,8,4,maxf=mf,twine=True)
