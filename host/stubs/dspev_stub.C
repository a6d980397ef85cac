// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for LAPACK dspev_ (no LAPACK in this image).
// The reference calls it at exactly one site (EW.C:5105, computeDT on the curvilinear grid)
// with N=3, JOBZ='N', packed lower-triangular storage {A11,A21,A31,A22,A32,A33}, and only
// reads the eigenvalues W (ascending).  Cyclic Jacobi rotations on the 3x3 matrix.
#include <cmath>
#include <algorithm>
extern "C" void dspev_( char& JOBZ, char& UPLO, int& N, double* AP, double* W, double* Z,
			int& LDZ, double* WORK, int& INFO )
{
   (void)JOBZ; (void)Z; (void)LDZ; (void)WORK;
   INFO = 0;
   if( N != 3 ) { INFO = -3; return; }
   double a[3][3];
   if( UPLO == 'L' || UPLO == 'l' )
   {
      a[0][0]=AP[0]; a[1][0]=AP[1]; a[2][0]=AP[2]; a[1][1]=AP[3]; a[2][1]=AP[4]; a[2][2]=AP[5];
   }
   else
   {
      a[0][0]=AP[0]; a[1][0]=AP[1]; a[1][1]=AP[2]; a[2][0]=AP[3]; a[2][1]=AP[4]; a[2][2]=AP[5];
   }
   a[0][1]=a[1][0]; a[0][2]=a[2][0]; a[1][2]=a[2][1];
   for( int sweep=0 ; sweep < 60 ; sweep++ )
   {
      double off = a[0][1]*a[0][1]+a[0][2]*a[0][2]+a[1][2]*a[1][2];
      double diag= a[0][0]*a[0][0]+a[1][1]*a[1][1]+a[2][2]*a[2][2];
      if( off <= 1e-32*diag || off == 0 ) break;
      for( int p=0 ; p < 2 ; p++ )
	 for( int q=p+1 ; q < 3 ; q++ )
	 {
	    if( a[p][q] == 0 ) continue;
	    double theta = (a[q][q]-a[p][p])/(2*a[p][q]);
	    double t = (theta >= 0 ? 1.0 : -1.0)/(fabs(theta)+sqrt(theta*theta+1));
	    double c = 1/sqrt(t*t+1), s = t*c;
	    for( int k=0 ; k < 3 ; k++ )
	    {
	       double akp = a[k][p], akq = a[k][q];
	       a[k][p] = c*akp - s*akq;
	       a[k][q] = s*akp + c*akq;
	    }
	    for( int k=0 ; k < 3 ; k++ )
	    {
	       double apk = a[p][k], aqk = a[q][k];
	       a[p][k] = c*apk - s*aqk;
	       a[q][k] = s*apk + c*aqk;
	    }
	 }
   }
   W[0]=a[0][0]; W[1]=a[1][1]; W[2]=a[2][2];
   std::sort( W, W+3 );
}
