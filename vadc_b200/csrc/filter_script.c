/* vadc_b200/csrc/filter_script.c -- timestamps -> ffmpeg audio filter script.
 *
 * Same contract as the reference's filter_script.c:4-23 (used by silenceremove.bat:11): "start,end" lines on stdin
 * (seconds, or centiseconds with --output_centi_seconds; both parse as floats) become
 *     asetpts=N/SR/TB, aselect='between(t,S,E)+between(t,S,E)+...', asetpts=N/SR/TB
 * on stdout, numbers printed with "%f", no trailing newline.
 * Differences: reading stops at the first line that is not a pair instead of spinning on it (the reference loops while
 * scanf_s() != EOF, which never ends on malformed input), and the multi-file output of vadc_b200_cli is understood: a
 * "# path" line starts a new script, which is printed after a copy of that line, one script per line.
 */
#include <stdio.h>
#include <string.h>

static void script_begin( void ) { fputs( "asetpts=N/SR/TB, aselect='", stdout ); }
static void script_end( void ) { fputs( "', asetpts=N/SR/TB", stdout ); }

int main( void )
{
   char line[512];
   int pairs = 0, scripts = 0, open = 0;
   while ( fgets( line, sizeof( line ), stdin ) )
   {
      if ( line[0] == '#' )
      {
         if ( open )
         {
            script_end();
            fputc( '\n', stdout );
         }
         fputs( line, stdout );
         if ( !strchr( line, '\n' ) ) fputc( '\n', stdout );
         script_begin();
         open = 1;
         pairs = 0;
         ++scripts;
         continue;
      }
      float from, to;
      if ( sscanf( line, "%f,%f", &from, &to ) != 2 )
      {
         if ( line[strspn( line, " \t\r\n" )] == 0 ) continue; /* blank line */
         break;
      }
      if ( !open )
      {
         script_begin();
         open = 1;
      }
      if ( pairs > 0 ) fputc( '+', stdout );
      fprintf( stdout, "between(t,%f,%f)", (double)from, (double)to );
      ++pairs;
   }
   if ( !open ) script_begin(); /* empty input still yields a (select-nothing) script, like the reference */
   script_end();
   if ( scripts ) fputc( '\n', stdout );
   return 0;
}
